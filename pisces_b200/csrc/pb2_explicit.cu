// Host side of the explicit-candidate path: the candidate table (RegionState.AddCandidate), VariantCollapser, MnvReallocator and the
// AlleleCaller.Call orchestration around the device kernels of pb2_candidates.cu / pb2_reads.cu. All counting, coverage and scoring runs
// on the GPU; what runs here is the reference's own host-side bookkeeping over a handful of candidates per block.
//
// Follows (reference @ /root/reference):
//   RegionState.AddCandidate / UpdateMaxPosition            src/lib/Pisces.Processing/RegionState/RegionState.cs:94-223
//   VariantCollapser.Collapse / CanCollapse / GetMatches     src/exe/Pisces/Logic/VariantCalling/VariantCollapser.cs:31-245
//   MnvReallocator.ReallocateFailedMnvs                      src/exe/Pisces/Logic/VariantCalling/MnvReallocator.cs:12-265
//   AlleleCaller.Call / GetRefSupportFromGappedMnvs          src/exe/Pisces/Logic/VariantCalling/AlleleCaller.cs:50-141,186-206
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <deque>
#include <memory>
#include <set>
#include <unordered_map>
#include <chrono>
#include "pb2_internal.hpp"

using namespace pb2;

#define CUX(h, expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) return pb2_fail((h), PB2_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ------------------------------------------------------------------------------------------------ the candidate table
static int block_key(int32_t position) { return (position + 999) / 1000; }   // RegionStateManager.GetBlockKey (:385-391), position >= 1

void explicit_add_candidate(pb2_handle* h, const HostCand& nc) {
    // UpdateMaxPosition (:204-223): never reset while the handle lives (RegionState.MaxAlleleEndpoint survives Reset())
    int other_end = 0;
    switch (nc.type) {
        case CAT_DEL: other_end = nc.position + (int)nc.ref.size(); break;
        case CAT_INS: other_end = nc.position + 1; break;
        case CAT_MNV: other_end = nc.position + (int)nc.ref.size() - 1; break;
        default: break;
    }
    int32_t& mx = h->block_max_endpoint[block_key(nc.position)];
    if (other_end > mx) mx = other_end;
    const bool track_open = h->cfg.collapse != 0;   // trackOpenEnded (Factory.cs:209-227)
    std::vector<uint32_t>& at = h->cand_by_pos[nc.position];   // the position's candidate list (RegionState._candidateVariantsLookup), in insertion order
    for (uint32_t idx : at) {
        HostCand& c = h->cands[idx];
        if (!c.alive) continue;
        if (c.Equals(nc) && (!track_open || (c.open_left == nc.open_left && c.open_right == nc.open_right))) {
            for (int i = 0; i < 3; i++) { c.support[i] += nc.support[i]; c.well_anchored[i] += nc.well_anchored[i]; }
            for (int i = 0; i < 8; i++) c.collapsed_mut[i] += nc.collapsed_mut[i];
            return;
        }
    }
    at.push_back((uint32_t)h->cands.size());
    h->cands.push_back(nc);
    h->cands.back().alive = true;
}
void explicit_reindex(pb2_handle* h) {   // after h->cands was compacted
    h->cand_by_pos.clear();
    for (size_t i = 0; i < h->cands.size(); i++) h->cand_by_pos[h->cands[i].position].push_back((uint32_t)i);
}

// ------------------------------------------------------------------------------------------------ device context of one batch
namespace {

// Device scratch of the explicit pass, from the handle's pool (stream-ordered: no cudaMalloc / cudaFree on the steady-state path).
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    pb2_handle* owner = nullptr;
    ~DevBuf() { if (p) { if (owner) pb2_dev_free(owner, p); else cudaFree(p); } }
    cudaError_t reserve(size_t n, cudaStream_t st, bool keep = false, pb2_handle* h = nullptr) {
        if (n <= cap) return cudaSuccess;
        const size_t ncap = std::max<size_t>(n, cap * 2 + 64);
        T* np = nullptr;
        cudaError_t e = h ? pb2_dev_alloc(h, reinterpret_cast<void**>(&np), ncap * sizeof(T)) : cudaMalloc(&np, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (keep && p && cap) e = cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (p) { if (owner) pb2_dev_free(owner, p); else { cudaStreamSynchronize(st); cudaFree(p); } }
        p = np; cap = ncap; owner = h;
        return e;
    }
};

// An allele in flight inside AlleleCaller.Call: CalledAllele's fields that the host logic touches.
struct Piece {
    int32_t position = 0;
    uint8_t type = 0;
    std::string ref, alt;
    int32_t allele_support = 0;
    int32_t support[3] = {0, 0, 0};
    int32_t well_anchored = 0;
    int32_t wa[3] = {0, 0, 0};                      // WellAnchoredSupportByDirection
    int32_t collapsed_mut[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // ReadCollapsedCountsMut (candidates only; alleles the reallocator creates start from zero)
    bool from_candidate = false;
    int32_t original_support = 0;   // reference alleles drawn into the MNV reallocation: re-scored only when their support changed
    bool keep_ref = false;          // the Reference candidate of a forced position (reference calls off): always scored and reported
};

struct BatchCtx {
    pb2_handle* h;
    std::unordered_map<int32_t, int32_t> row_of_pos;   // reference position -> row of the gathered tables (-1: position not staged)
    int32_t n_rows = 0;
    DevBuf<int32_t> counts, collapsed, req;
    DevBuf<double> qsum;
    DevBuf<DevCand> d_cands;
    DevBuf<uint8_t> d_arena, d_flags;
    DevBuf<pb2_call_record> d_out;
    DevBuf<SpanIngredients> d_ingr;
    bool want_q;
    bool use_resident_reads = false;   // pb2_call_resident: the segment pb2_stage_reads built is the one to gather from
    struct GatherRec { size_t seg; int32_t req_off, n, row0; };
    std::vector<GatherRec> gathers;   // the gather launches made so far (replayed by pb2_call_resident)
    explicit BatchCtx(pb2_handle* hh) : h(hh) { want_q = hh->cfg.want_sum_base_quality || hh->cfg.noise_model == 1; }

    // locus of a reference position in a staged segment
    static int64_t locus_of(const Segment& s, int32_t pos) {
        if (s.has_positions) {
            auto it = std::lower_bound(s.h_positions.begin(), s.h_positions.end(), pos);
            return (it != s.h_positions.end() && *it == pos) ? (int64_t)(it - s.h_positions.begin()) : -1;
        }
        const int64_t k = (int64_t)pos - s.first_position;
        return (k >= 0 && k < s.n_loci) ? k : -1;
    }
    static TilePileup view(const Segment& s) {
        TilePileup in;
        memset(&in, 0, sizeof(in));
        in.cq = s.code; in.anch = s.anch; in.tile_base = s.tile_base; in.depth = s.depth; in.pad = s.pad; in.ref_base = s.ref_base;
        in.positions = s.positions; in.first_position = s.first_position; in.n_loci = s.n_loci; in.n_tiles = s.n_tiles;
        in.plane_bytes = std::max<int64_t>(s.plane_bytes, 16);
        return in;
    }
    // 198-bin counts of the requested loci of a segment, whichever form it is staged in
    static cudaError_t gather_rows(const Segment& s, const int32_t* req, int32_t n, int32_t* counts, int32_t* collapsed, double* qsum, int min_bq, cudaStream_t st) {
        if (s.pv_data != nullptr) return launch_pvert_gather(pvert_view(s), req, n, counts, collapsed, min_bq, st);   // (no quality sums in this form: pvert_eligible)
        return launch_gather_locus_counts(view(s), req, n, counts, collapsed, qsum, min_bq, st);
    }
    // make sure the count tables hold a row for each of these positions (gathered from whichever segment stages it)
    int ensure_rows(const std::vector<int32_t>& positions) {
        std::vector<std::vector<std::pair<int32_t, int32_t>>> per_seg(h->segs.size());   // (row, locus)
        int32_t next = n_rows;
        for (int32_t pos : positions) {
            if (row_of_pos.count(pos)) continue;
            int32_t row = -1;
            for (size_t si = 0; si < h->segs.size(); si++) {
                if (h->segs[si].from_reads && !use_resident_reads) continue;
                const int64_t l = locus_of(h->segs[si], pos);
                if (l >= 0) { row = next++; per_seg[si].push_back({row, (int32_t)l}); break; }
            }
            row_of_pos[pos] = row;   // -1: the reference reads 0 from a block that does not exist (RegionStateManager.cs:224-225)
        }
        if (next == n_rows) return PB2_OK;
        cudaStream_t st = h->stream;
        CUX(h, counts.reserve((size_t)next * kNumBins, st, true, h));
        CUX(h, collapsed.reserve((size_t)next * kNumCollapsed, st, true, h));
        if (want_q) CUX(h, qsum.reserve((size_t)next * kNumBins, st, true, h));
        // rows were numbered segment by segment in position order within this call; launch one gather per segment over its contiguous run
        std::vector<int32_t> loci((size_t)(next - n_rows), -1);
        for (auto& v : per_seg) for (auto& rl : v) loci[(size_t)(rl.first - n_rows)] = rl.second;
        CUX(h, req.reserve(loci.size(), st, false, h));
        CUX(h, cudaMemcpyAsync(req.p, loci.data(), loci.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        for (size_t si = 0; si < h->segs.size(); si++) {
            if (per_seg[si].empty()) continue;
            // rows of one segment are not necessarily contiguous (positions interleave across segments): gather them run by run
            size_t a = 0;
            auto& v = per_seg[si];
            while (a < v.size()) {
                size_t b = a + 1;
                while (b < v.size() && v[b].first == v[b - 1].first + 1) b++;
                const int32_t r0 = v[a].first, n = (int32_t)(b - a);
                CUX(h, gather_rows(h->segs[si], req.p + (r0 - n_rows), n, counts.p + (size_t)r0 * kNumBins, collapsed.p + (size_t)r0 * kNumCollapsed,
                                   want_q ? qsum.p + (size_t)r0 * kNumBins : nullptr, h->dcfg.min_bq, st));
                h->total_launches += 1;
                gathers.push_back({si, r0 - n_rows, n, r0});
                a = b;
            }
        }
        CUX(h, cudaStreamSynchronize(st));   // req is reused by the next call
        n_rows = next;
        return PB2_OK;
    }
};

bool in_intervals(const pb2_handle* h, int32_t pos) {   // AlleleCaller.ShouldReport (:260-263)
    if (!h->have_intervals) return true;
    for (size_t i = 0; i < h->iv_start.size(); i++) if (pos >= h->iv_start[i] && pos <= h->iv_end[i]) return true;
    return false;
}

void end_points(uint8_t type, int32_t position, int ref_len, int alt_len, int32_t& start, int32_t& end) {   // CoverageCalculator.Compute (:19-47)
    switch (type) {
        case CAT_DEL: start = position + 1; end = position + (ref_len - 1); break;
        case CAT_MNV: start = position; end = position + alt_len - 1; break;
        case CAT_INS: start = position; end = position + 1; break;
        default: start = position; end = position; break;
    }
}

// Scores `pieces` on the device. Dense mode: records / flags / ingredients come back to the host. Append mode (seg != nullptr): callable alleles are
// appended to the segment's variant stream and prune its reference records; nothing comes back.
int score_pieces(BatchCtx& ctx, const std::vector<Piece*>& pieces, std::vector<uint8_t>& arena, std::vector<pb2_call_record>* out_records,
                 std::vector<uint8_t>* out_flags, std::vector<SpanIngredients>* out_ingr, Segment* append_seg) {
    pb2_handle* h = ctx.h;
    const size_t n = pieces.size();
    if (n == 0) return PB2_OK;
    std::vector<int32_t> need;
    for (auto* p : pieces) {
        int32_t s, e;
        end_points(p->type, p->position, (int)p->ref.size(), (int)p->alt.size(), s, e);
        need.push_back(s);
        if (e != s) need.push_back(e);
    }
    std::sort(need.begin(), need.end());
    need.erase(std::unique(need.begin(), need.end()), need.end());
    int rc = ctx.ensure_rows(need);
    if (rc != PB2_OK) return rc;
    std::vector<DevCand> dc(n);
    for (size_t i = 0; i < n; i++) {
        const Piece& p = *pieces[i];
        DevCand& d = dc[i];
        memset(&d, 0, sizeof(d));
        d.position = p.position;
        d.type = p.type;
        d.flags = (p.alt.find('N') != std::string::npos ? kCandAltHasN : 0) | (in_intervals(h, p.position) ? kCandReportable : 0) |
                  (!h->forced.empty() && h->forced.count(std::make_tuple(p.position, p.ref, p.alt)) ? kCandForced : 0);
        d.ref_len = (int32_t)p.ref.size();
        d.alt_len = (int32_t)p.alt.size();
        d.allele_off = (uint32_t)arena.size();
        arena.insert(arena.end(), p.ref.begin(), p.ref.end());
        arena.insert(arena.end(), p.alt.begin(), p.alt.end());
        // AlleleSupport and SupportByDirection move together everywhere in the reference (MnvReallocator.ProcessOverlap :100-112)
        for (int k = 0; k < 3; k++) d.support[k] = p.support[k];
        d.well_anchored = p.well_anchored;
        int32_t s, e;
        end_points(p.type, p.position, d.ref_len, d.alt_len, s, e);
        d.req_start = ctx.row_of_pos[s];
        d.req_end = ctx.row_of_pos[e];
        auto g = h->gapped_ref.find(p.position);
        d.gapped_ref = g == h->gapped_ref.end() ? 0 : g->second;
        d.locus = append_seg ? (int32_t)BatchCtx::locus_of(*append_seg, p.position) : -1;
    }
    cudaStream_t st = h->stream;
    CUX(h, ctx.d_cands.reserve(n, st, false, h));
    CUX(h, ctx.d_arena.reserve(arena.size() + 16, st, false, h));
    CUX(h, cudaMemcpyAsync(ctx.d_cands.p, dc.data(), n * sizeof(DevCand), cudaMemcpyHostToDevice, st));
    CUX(h, cudaMemcpyAsync(ctx.d_arena.p, arena.data(), arena.size(), cudaMemcpyHostToDevice, st));
    CandScoreArgs a;
    memset(&a, 0, sizeof(a));
    a.cands = ctx.d_cands.p; a.n = (int32_t)n; a.counts = ctx.counts.p; a.collapsed = ctx.collapsed.p; a.qsum = ctx.want_q ? ctx.qsum.p : nullptr;
    a.arena = ctx.d_arena.p; a.chr_seq = h->d_chr; a.chr_len = h->chr_len; a.q_to_p_table = h->d_q_to_p; a.q_table_max = h->q_table_max;
    a.indel_repeat_filter = h->cfg.indel_repeat_filter;
    if (append_seg) {
        a.var_records = append_seg->var_records; a.var_count = append_seg->counters; a.var_capacity = append_seg->var_capacity; a.ref_valid = append_seg->ref_valid;
    } else {
        CUX(h, ctx.d_out.reserve(n, st, false, h));
        CUX(h, ctx.d_flags.reserve(n, st, false, h));
        CUX(h, ctx.d_ingr.reserve(n, st, false, h));
        a.out_dense = ctx.d_out.p; a.out_callable = ctx.d_flags.p; a.out_ingredients = out_ingr ? ctx.d_ingr.p : nullptr;
    }
    CUX(h, launch_score_candidates(a, h->dcfg, st));
    h->total_launches += 1;
    // AmpliconBiasCalculator.Compute for the SNV candidates among them (positions whose SNVs were made explicit): per-amplicon tallies from the pileup of
    // the segment that holds the position. (In append mode the pass over the segment's variant stream covers them: amplicon_pass, pb2_api.cu.)
    int amp_status = 0;
    if (!append_seg && h->cfg.amplicon_bias_filter >= 0) {
        for (auto& seg : h->segs) {
            if (seg.pv_row_amp == nullptr) continue;
            int* d_status = reinterpret_cast<int*>(seg.counters + 4);
            CUX(h, cudaMemsetAsync(d_status, 0, sizeof(int), st));
            CUX(h, launch_pvert_amplicon_bias(pvert_view(seg), seg.pv_row_amp, ctx.d_out.p, nullptr, (int64_t)n, (int64_t)n, nullptr, h->cfg.amplicon_bias_filter, h->dcfg.min_bq, d_status,
                                              h->num_sms, st));
            int one = 0;
            CUX(h, cudaMemcpyAsync(&one, d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUX(h, cudaStreamSynchronize(st));
            amp_status |= one;
            h->total_launches += 1;
        }
    }
    if (amp_status) { h->error = "Index was outside the bounds of the array."; return PB2_ERR_ARG; }
    if (!append_seg) {
        if (out_records) { out_records->resize(n); CUX(h, cudaMemcpyAsync(out_records->data(), ctx.d_out.p, n * sizeof(pb2_call_record), cudaMemcpyDeviceToHost, st)); }
        if (out_flags) { out_flags->resize(n); CUX(h, cudaMemcpyAsync(out_flags->data(), ctx.d_flags.p, n, cudaMemcpyDeviceToHost, st)); }
        if (out_ingr) { out_ingr->resize(n); CUX(h, cudaMemcpyAsync(out_ingr->data(), ctx.d_ingr.p, n * sizeof(SpanIngredients), cudaMemcpyDeviceToHost, st)); }
    }
    CUX(h, cudaStreamSynchronize(st));
    return PB2_OK;
}

Piece piece_of(const HostCand& c) {   // AlleleHelper.Map (Utility/AlleleHelper.cs:34-60)
    Piece p;
    p.position = c.position; p.type = c.type; p.ref = c.ref; p.alt = c.alt;
    p.allele_support = c.Support();
    for (int k = 0; k < 3; k++) p.support[k] = c.support[k];
    p.well_anchored = c.WellAnchored();
    for (int k = 0; k < 3; k++) p.wa[k] = c.well_anchored[k];
    for (int k = 0; k < 8; k++) p.collapsed_mut[k] = c.collapsed_mut[k];
    p.from_candidate = true;
    return p;
}

// ------------------------------------------------------------------------------------------------ VariantCollapser
bool can_collapse(const HostCand& tc, const HostCand& pm) {   // VariantCollapser.CanCollapse (:125-175)
    if ((tc.type == CAT_INS) != (pm.type == CAT_INS) || (tc.type == CAT_DEL) != (pm.type == CAT_DEL) || tc.Length() > pm.Length() ||
        (tc.FullyAnchored() && !pm.FullyAnchored()))
        return false;
    const std::string& tcb = tc.type == CAT_DEL ? tc.ref : tc.alt;
    const std::string& pmb = pm.type == CAT_DEL ? pm.ref : pm.alt;
    if (tc.FullyAnchored() && pm.FullyAnchored()) return tc.Equals(pm);
    if (tc.type == CAT_DEL) {
        if (tc.open_right) return pm.position + 1 == tc.position + 1;
        return pm.position + (int)pmb.size() - 1 == tc.position + (int)tcb.size() - 1;
    }
    if (tc.open_right) return pm.position == tc.position && pmb.compare(0, tcb.size(), tcb) == 0;
    if (tc.type == CAT_INS)
        return pm.position + 1 == tc.position + 1 && pmb.size() + 1 >= tcb.size() && pmb.substr(pmb.size() - tcb.size() + 1) == tcb.substr(1);
    return pm.position + (int)pm.alt.size() - 1 == tc.position + (int)tc.alt.size() - 1 && pm.alt.size() >= tc.alt.size() &&
           pm.alt.compare(pm.alt.size() - tc.alt.size(), tc.alt.size(), tc.alt) == 0;
}
int compare_matches(const HostCand& a, const HostCand& b) {   // VariantCollapser.Compare (:221-245); IsKnown is never set (no priors)
    if (a.FullyAnchored() && !b.FullyAnchored()) return -1;
    if (!a.FullyAnchored() && b.FullyAnchored()) return 1;
    if (a.Length() != b.Length()) return a.Length() < b.Length() ? 1 : -1;
    if (std::fabs(a.frequency - b.frequency) > 0.0f) return a.frequency < b.frequency ? 1 : -1;
    if (a.position != b.position) return a.position < b.position ? -1 : 1;
    const int c = a.alt.compare(b.alt);
    return c < 0 ? -1 : (c > 0 ? 1 : 0);
}
// CalledAllele.Frequency after CoverageCalculator.Compute for a candidate with its current support
float candidate_frequency(const pb2_handle* h, const HostCand& c, const SpanIngredients& g, int point_total) {
    int total = point_total;
    if (c.type == CAT_INS || c.type == CAT_DEL || c.type == CAT_MNV) {
        const bool presume = c.type == CAT_INS ? (h->cfg.expect_stitched != 0) : true;
        total = spanning_tail(g, c.type == CAT_INS, presume, c.Support(), c.WellAnchored()).total;
    }
    if (total == 0) return 0.0f;
    const float f = (float)c.Support() / (float)total;
    return f < 1.0f ? f : 1.0f;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ forced-genotyping alleles
int explicit_add_forced_candidates(pb2_handle* h, int32_t up_to) {   // SmallVariantCaller.AddForcedAlleleAsCandidate (:118-155)
    while (!h->forced_pending.empty()) {
        auto it = h->forced_pending.begin();
        if (up_to >= 0 && it->first > up_to) break;
        for (auto& ra : it->second) {
            HostCand c;
            c.position = it->first; c.ref = ra.first; c.alt = ra.second;
            // CandidateAllele type of a forced allele (:131-147)
            c.type = (c.ref.size() == 1 && c.alt.size() == 1) ? CAT_SNV : c.ref.size() == c.alt.size() ? CAT_MNV : c.ref.size() > c.alt.size() ? CAT_DEL : CAT_INS;
            explicit_add_candidate(h, c);
        }
        h->forced_pending.erase(it);
    }
    return PB2_OK;
}

// ------------------------------------------------------------------------------------------------ AlleleCaller.Call, explicit part
// PB2_TRACE: where the time of the batches goes (host phases of explicit_call_batch, summed over the batches of a flush)
double g_batch_phase_ms[8];
static bool batch_trace() { static const bool on = getenv("PB2_TRACE") != nullptr; return on; }
struct PhaseClock {
    std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    void mark(int phase) {
        if (!batch_trace()) return;
        const auto now = std::chrono::steady_clock::now();
        g_batch_phase_ms[phase] += std::chrono::duration<double, std::milli>(now - last).count();
        last = now;
    }
};
int explicit_call_batch(pb2_handle* h, const std::vector<size_t>& batch, int32_t max_cleared, int32_t ref_lo, int32_t ref_hi, std::vector<pb2_call_record>& called,
                        std::vector<pb2_call_record_ext>& called_ext, const std::vector<size_t>* kill) {
    if (batch.empty()) return PB2_OK;
    PhaseClock pc_;
    std::vector<HostCand> cs;
    cs.reserve(batch.size());
    for (size_t idx : batch) { cs.push_back(h->cands[idx]); cs.back().alive = true; }
    for (size_t idx : (kill ? *kill : batch)) h->cands[idx].alive = false;   // the batch leaves the state (DoneProcessing / ExtractCollapsable)
    BatchCtx ctx(h);
    std::vector<uint8_t>& arena = h->arena;
    int rc;
    // anchor-summed counts [allele][direction] of a few positions on the host (what RegionState.GetAllCandidates :417-440 sums for a reference candidate)
    auto host_point_counts = [&](std::vector<int32_t> positions, std::unordered_map<int32_t, std::array<int32_t, kNumAlleles * kNumDirs>>& out) -> int {
        std::sort(positions.begin(), positions.end());
        positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
        const int r0 = ctx.ensure_rows(positions);
        if (r0 != PB2_OK) return r0;
        std::vector<int32_t> table((size_t)ctx.n_rows * kNumBins);
        if (!table.empty()) CUX(h, cudaMemcpy(table.data(), ctx.counts.p, table.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (int32_t pos : positions) {
            std::array<int32_t, kNumAlleles * kNumDirs> sums{};
            const int32_t r = ctx.row_of_pos[pos];
            if (r >= 0)
                for (int ad = 0; ad < kNumAlleles * kNumDirs; ad++)
                    for (int an = 0; an < kNumAnchors; an++) sums[(size_t)ad] += table[(size_t)r * kNumBins + (size_t)(ad * kNumAnchors + an)];
            out[pos] = sums;
        }
        return PB2_OK;
    };
    auto allele_index = [](char base) { return base == 'A' ? AT_A : base == 'C' ? AT_C : base == 'G' ? AT_G : base == 'T' ? AT_T : AT_N; };
    if (!h->cfg.call_mnvs) {
        // CallMNVs off: SNV candidates are the counts themselves and never enter this table — except forced SNV alleles, which arrive with zero
        // support and merge with what the finder raised from the reads (RegionState.AddCandidate :94-174): that support is the count of their base
        // (positions whose SNV candidates were made explicit carry the finder's own support: explicit_materialize_snvs)
        auto count_based = [&](const HostCand& c) {
            if (c.type != CAT_SNV) return false;
            for (auto& r : h->snv_explicit_ranges) if (c.position > r.first && c.position <= r.second) return false;
            return true;
        };
        std::vector<int32_t> snv_pos;
        for (auto& c : cs) if (count_based(c)) snv_pos.push_back(c.position);
        if (!snv_pos.empty()) {
            std::unordered_map<int32_t, std::array<int32_t, kNumAlleles * kNumDirs>> pc;
            rc = host_point_counts(snv_pos, pc);
            if (rc != PB2_OK) return rc;
            for (auto& c : cs)
                if (count_based(c)) for (int d = 0; d < 3; d++) c.support[d] = pc[c.position][(size_t)(allele_index(c.alt[0]) * kNumDirs + d)];
        }
    }

    pc_.mark(0);
    // ---- VariantCollapser.Collapse (:31-113)
    const bool any_open = std::any_of(cs.begin(), cs.end(), [](const HostCand& c) { return c.open_left || c.open_right; });
    if (h->cfg.collapse && any_open) {
        // coverage ingredients of every candidate, once: the frequencies the collapser compares are re-derived from them after each merge
        std::deque<Piece> store;
        std::vector<Piece*> ps;
        for (auto& c : cs) { store.push_back(piece_of(c)); ps.push_back(&store.back()); }
        std::vector<pb2_call_record> recs;
        std::vector<SpanIngredients> ingr;
        std::vector<uint8_t> scratch_arena;
        rc = score_pieces(ctx, ps, scratch_arena, &recs, nullptr, &ingr, nullptr);
        if (rc != PB2_OK) return rc;
        pc_.mark(1);
        std::vector<size_t> targets;
        for (size_t i = 0; i < cs.size(); i++) if (!(h->cfg.exclude_mnvs_from_collapsing && cs[i].type == CAT_MNV)) targets.push_back(i);
        std::vector<size_t> to_collapse;
        for (size_t i : targets) if (cs[i].open_left || cs[i].open_right) to_collapse.push_back(i);
        std::stable_sort(to_collapse.begin(), to_collapse.end(), [&](size_t ia, size_t ib) {   // the LINQ OrderBy chain (:42-47)
            const HostCand &a = cs[ia], &b = cs[ib];
            if (a.Length() != b.Length()) return a.Length() > b.Length();
            const bool ab = a.open_left && a.open_right, bb = b.open_left && b.open_right;
            if (ab != bb) return ab;
            const bool ao = a.open_left || a.open_right, bo = b.open_left || b.open_right;
            if (ao != bo) return ao;
            if (a.ref != b.ref) return a.ref < b.ref;
            if (a.alt != b.alt) return a.alt < b.alt;
            if (a.Support() != b.Support()) return a.Support() < b.Support();
            if (a.open_right != b.open_right) return !a.open_right;
            if (a.open_left != b.open_left) return !a.open_left;
            return false;
        });
        // A target can only take a candidate that starts or ends where it does (can_collapse compares one of the two, by type and open end), so the
        // targets are indexed by both positions once per batch: with CallMNVs every sequencing error is a candidate, thousands per block, and the
        // all-pairs scan was the flush's largest item
        auto end_of = [](const HostCand& c) { return c.position + (int32_t)(c.type == CAT_DEL ? c.ref.size() : c.alt.size()) - 1; };
        std::unordered_map<int32_t, std::vector<size_t>> by_start, by_end;
        by_start.reserve(targets.size()); by_end.reserve(targets.size());
        for (size_t t : targets) { by_start[cs[t].position].push_back(t); by_end[end_of(cs[t])].push_back(t); }
        std::vector<size_t> near;
        for (size_t vi : to_collapse) {
            HostCand& v = cs[vi];
            near.clear();
            auto a = by_start.find(v.position);
            if (a != by_start.end()) near.insert(near.end(), a->second.begin(), a->second.end());
            auto b = by_end.find(end_of(v));
            if (b != by_end.end()) near.insert(near.end(), b->second.begin(), b->second.end());
            std::sort(near.begin(), near.end());   // the order of `targets` (= candidate order), each target once
            near.erase(std::unique(near.begin(), near.end()), near.end());
            std::vector<size_t> pms;
            for (size_t t : near) if (t != vi && cs[t].alive && can_collapse(v, cs[t])) pms.push_back(t);
            if (pms.empty()) continue;
            for (size_t t : pms) cs[t].frequency = candidate_frequency(h, cs[t], ingr[t], recs[t].total_coverage);
            const float tcf = candidate_frequency(h, v, ingr[vi], recs[vi].total_coverage);
            std::stable_sort(pms.begin(), pms.end(), [&](size_t a, size_t b) { return compare_matches(cs[a], cs[b]) < 0; });
            long match = -1;
            for (size_t t : pms) if (cs[t].Equals(v) && !cs[t].open_left && !cs[t].open_right) { match = (long)t; break; }
            if (match < 0)
                for (size_t t : pms)
                    if (cs[t].frequency >= h->cfg.collapse_freq_threshold && cs[t].frequency / tcf > h->cfg.collapse_freq_ratio_threshold) { match = (long)t; break; }
            if (match < 0) continue;
            HostCand& m = cs[(size_t)match];
            h->total_collapsed++;
            for (int k = 0; k < 3; k++) { m.support[k] += v.support[k]; m.well_anchored[k] += v.well_anchored[k]; }
            m.open_left = m.open_left && v.open_left;
            m.open_right = m.open_right && v.open_right;
            for (int k = 0; k < 8; k++) m.collapsed_mut[k] += v.collapsed_mut[k];
            v.alive = false;
        }
    }
    pc_.mark(2);
    if (h->cfg.collapse && max_cleared >= 0) {   // candidates beyond the cleared positions go back to the state (:100-112)
        for (auto& c : cs)
            if (c.alive && c.position > max_cleared) { c.alive = false; HostCand back = c; back.alive = true; explicit_add_candidate(h, back); }
    }

    // ---- MNVs first: ProcessVariant + IsCallable; the failed ones are reallocated (AlleleCaller.cs:62-92)
    std::deque<Piece> store;
    std::vector<Piece*> callable, failed;
    {
        std::vector<Piece*> mnvs;
        std::vector<Piece*> in_order;
        for (auto& c : cs) {
            if (!c.alive) continue;
            store.push_back(piece_of(c));
            in_order.push_back(&store.back());
            if (c.type == CAT_MNV) mnvs.push_back(&store.back());
        }
        std::vector<uint8_t> flags;
        if (!mnvs.empty()) {
            std::vector<uint8_t> scratch_arena;
            rc = score_pieces(ctx, mnvs, scratch_arena, nullptr, &flags, nullptr, nullptr);
            if (rc != PB2_OK) return rc;
        }
        size_t mi = 0;
        for (Piece* p : in_order) {
            if (p->type == CAT_MNV) { if (flags[mi++] & 1) callable.push_back(p); else failed.push_back(p); }
            else callable.push_back(p);
        }
    }
    pc_.mark(3);
    {
        // Reference candidates of the batch (RegionState.GetAllCandidates :393-449). With reference calls on there is one per position and the
        // reallocator treats them like any callable allele (IsPotentialOverlap :262): a failed gapped MNV hands its support to the reference allele at a
        // position where its alternate base equals the reference base; only those positions can match (OverlapMatches), so only they are built here —
        // the rest stay with the hot kernel's per-locus reference stream. With reference calls off and forced alleles present there is one at every
        // forced position of the batch's blocks, reported whatever its support.
        std::vector<int32_t> gap_pos, forced_pos;
        if (!failed.empty() && h->cfg.output_gvcf)
            for (Piece* f : failed)
                for (size_t i = 0; i < f->ref.size(); i++) {
                    const int32_t pos = f->position + (int32_t)i;
                    if (f->ref[i] == f->alt[i] && pos >= 1 && pos <= h->chr_len && (max_cleared < 0 || pos <= max_cleared)) gap_pos.push_back(pos);
                }
        if (!h->cfg.output_gvcf && !h->forced_positions.empty()) {
            // block by block; inside a block ChrIntervalSet.GetClipped (IntervalSet.cs:76-105) walks the one-position intervals in the forced set's own
            // order and stops at the first one beyond the block — with an unsorted set that skips the rest, as in the reference
            std::set<int32_t> block_keys;
            for (int32_t pos : h->forced_positions) if (pos > ref_lo && pos <= ref_hi) block_keys.insert(block_key(pos));
            for (int32_t k : block_keys) {
                const int32_t start = (k - 1) * 1000 + 1, end = k * 1000;
                for (int32_t pos : h->forced_positions) {
                    if (pos > end) break;
                    if (pos < start) continue;
                    if (pos >= 1 && pos <= h->chr_len) forced_pos.push_back(pos);
                }
            }
        }
        std::sort(gap_pos.begin(), gap_pos.end());
        gap_pos.erase(std::unique(gap_pos.begin(), gap_pos.end()), gap_pos.end());
        std::vector<int32_t> all_pos = gap_pos;
        all_pos.insert(all_pos.end(), forced_pos.begin(), forced_pos.end());
        if (!all_pos.empty()) {
            std::unordered_map<int32_t, std::array<int32_t, kNumAlleles * kNumDirs>> pc;
            rc = host_point_counts(all_pos, pc);
            if (rc != PB2_OK) return rc;
            auto add_ref = [&](int32_t pos, bool forced_position) {
                const auto& sums = pc[pos];
                const char base = (char)h->h_chr[(size_t)pos - 1];
                const int ref_idx = allele_index(base);
                int total = 0;
                for (int32_t v : sums) total += v;
                if (!forced_position && !(h->have_intervals ? in_intervals(h, pos) : total > 0)) return;
                store.emplace_back();
                Piece* p = &store.back();
                p->position = pos; p->type = CAT_REF; p->ref.assign(1, base); p->alt.assign(1, base);
                for (int d = 0; d < 3; d++) p->support[d] = sums[(size_t)(ref_idx * kNumDirs + d)];
                p->allele_support = p->original_support = p->support[0] + p->support[1] + p->support[2];
                p->keep_ref = forced_position;
                callable.push_back(p);
            };
            for (int32_t pos : gap_pos) add_ref(pos, false);
            for (int32_t pos : forced_pos) add_ref(pos, true);
        }
    }
    pc_.mark(4);
    if (!failed.empty()) {
        // ---- MnvReallocator.ReallocateFailedMnvs (:12-98)
        auto create = [&](int32_t pos, int allele_support, const std::string& alt, const std::string& ref, const int32_t* sup) -> Piece* {   // CreateVariant (:156-173)
            store.emplace_back();
            Piece* a = &store.back();
            a->type = alt == ref ? CAT_REF : (alt.size() > 1 ? CAT_MNV : CAT_SNV);
            a->position = pos; a->allele_support = allele_support; a->alt = alt; a->ref = ref;
            if (sup) for (int k = 0; k < 3; k++) a->support[k] = sup[k];
            return a;
        };
        auto break_off = [&](Piece* al) -> Piece* {   // BreakOffEdgeReferences (:215-246)
            if (al->type != CAT_MNV) return al;
            int left = 0, right = 0;
            const int n = (int)al->ref.size();
            for (int i = 0; i < n; i++) { if (al->ref[(size_t)i] != al->alt[(size_t)i]) break; left++; }
            for (int i = 0; i < n; i++) { const int k = n - 1 - i; if (al->ref[(size_t)k] != al->alt[(size_t)k]) break; right++; }
            return create(al->position + left, al->allele_support, al->alt.substr((size_t)left, al->alt.size() - (size_t)(left + right)),
                          al->ref.substr((size_t)left, al->ref.size() - (size_t)(left + right)), al->support);
        };
        auto remove_ref = [](std::vector<Piece*>& v, Piece* x) { auto it = std::find(v.begin(), v.end(), x); if (it != v.end()) v.erase(it); };
        auto order_key = [](const Piece* a, const Piece* b) {   // OrderByDescending(alt length).ThenByDescending(support).ThenBy(alt).ThenBy(ref)
            if (a->alt.size() != b->alt.size()) return a->alt.size() > b->alt.size();
            if (a->allele_support != b->allele_support) return a->allele_support > b->allele_support;
            if (a->alt != b->alt) return a->alt < b->alt;
            return a->ref < b->ref;
        };
        const bool have_max = max_cleared >= 0;
        std::vector<Piece*> outside;
        auto process_overlap = [&](Piece* overlap, Piece* re, std::vector<Piece*>& remainder) {   // ProcessOverlap (:100-137)
            overlap->allele_support += re->allele_support;
            for (int k = 0; k < 3; k++) overlap->support[k] += re->support[k];
            remove_ref(remainder, re);
            // CreateAllelesFromRemainder (:175-213)
            std::vector<Piece*> rems;
            const int idx = overlap->position - re->position;
            const int right_side = idx + (int)overlap->alt.size();
            const int alt_len = (int)re->alt.size();
            if (alt_len - right_side > 0 && right_side <= re->position + alt_len) {
                Piece* rr = create(re->position + right_side, re->allele_support, re->alt.substr((size_t)right_side), re->ref.substr((size_t)right_side, (size_t)(alt_len - right_side)),
                                   re->support);
                if (rr->type != CAT_REF) rems.push_back(rr);
            }
            if (idx > 0) {
                Piece* lr = create(re->position, re->allele_support, re->alt.substr(0, (size_t)idx), re->ref.substr(0, (size_t)idx), re->support);
                if (lr->type != CAT_REF) rems.push_back(lr);
            }
            for (auto& r : rems) r = break_off(r);
            if (have_max) {
                if (overlap->position > max_cleared) { remove_ref(remainder, overlap); outside.push_back(overlap); }
                for (Piece* r : rems) { if (r->position <= max_cleared) remainder.push_back(r); else outside.push_back(r); }
            } else remainder.insert(remainder.end(), rems.begin(), rems.end());
        };
        std::vector<Piece*> ordered = failed;
        std::stable_sort(ordered.begin(), ordered.end(), [&](const Piece* a, const Piece* b) {
            if (a->position != b->position) return a->position < b->position;
            return order_key(a, b);
        });
        for (Piece* fm : ordered) {
            std::vector<Piece*> remainder{fm};
            while (!remainder.empty()) {
                Piece* re = remainder.front();
                std::vector<Piece*> overlaps;
                const int re_end = re->position + (int)re->alt.size();
                for (Piece* c : callable)   // IsPotentialOverlap (:255-265)
                    if (c->position >= re->position && c->position <= re_end && c->alt.size() <= re->alt.size() && c->position + (int)c->alt.size() <= re_end &&
                        (c->type == CAT_MNV || c->type == CAT_SNV || c->type == CAT_REF))
                        overlaps.push_back(c);
                std::stable_sort(overlaps.begin(), overlaps.end(), order_key);
                std::vector<Piece*> matching;
                for (Piece* o : overlaps)   // OverlapMatches (:248-253)
                    if (re->alt.compare((size_t)(o->position - re->position), o->alt.size(), o->alt) == 0) matching.push_back(o);
                bool reallocated = false;
                if (have_max) {
                    const int into_next = re->position + ((int)re->alt.size() - 1) - max_cleared;
                    const bool any_long = std::any_of(matching.begin(), matching.end(), [](const Piece* o) { return o->alt.size() > 1; });
                    if (into_next > 0 && !any_long) {
                        if (re->position <= max_cleared) {
                            const int orig_len = (int)re->ref.size();
                            Piece* nb = create(max_cleared + 1, 0, re->alt.substr((size_t)(orig_len - into_next), (size_t)into_next),
                                               re->ref.substr((size_t)(orig_len - into_next), (size_t)into_next), nullptr);
                            process_overlap(break_off(nb), re, remainder);
                        } else {
                            remove_ref(remainder, re);
                            outside.push_back(re);
                        }
                        reallocated = true;
                    }
                }
                if (!reallocated && !matching.empty()) { process_overlap(matching.front(), re, remainder); reallocated = true; }
                if (!reallocated) {   // BreakDownToSingleNucCalls (:139-154)
                    for (size_t i = 0; i < re->alt.size(); i++) {
                        Piece* sn = create(re->position + (int)i, re->allele_support, re->alt.substr(i, 1), re->ref.substr(i, 1), re->support);
                        if (sn->type == CAT_REF) continue;
                        if (have_max && !(sn->position <= max_cleared)) outside.push_back(sn);
                        else callable.push_back(sn);
                    }
                    remove_ref(remainder, re);
                }
            }
        }
        for (Piece* l : outside) {   // leftovers become candidates of later blocks (AlleleCaller.cs:91-92, AlleleHelper.Map :62-85)
            HostCand c;
            c.position = l->position; c.type = l->type; c.ref = l->ref; c.alt = l->alt;
            for (int k = 0; k < 3; k++) { c.support[k] = l->support[k]; c.well_anchored[k] = l->wa[k]; }
            for (int k = 0; k < 8; k++) c.collapsed_mut[k] = l->collapsed_mut[k];
            explicit_add_candidate(h, c);
        }
    }
    pc_.mark(5);
    // ---- GetRefSupportFromGappedMnvs (:186-206) -> RegionState.AddGappedMnvRefCount
    for (Piece* a : callable) {
        if (a->type != CAT_MNV) continue;
        for (size_t i = 0; i < a->ref.size(); i++)
            if (a->ref[i] == a->alt[i]) h->gapped_ref[a->position + (int)i] += a->allele_support;
    }
    // a failed MNV that is a forced allele is reported all the same (:98-106)
    if (!h->forced.empty())
        for (Piece* f : failed) if (h->forced.count(std::make_tuple(f->position, f->ref, f->alt))) callable.push_back(f);
    // reference alleles the reallocation left untouched stay with the per-locus reference stream of the hot kernel
    callable.erase(std::remove_if(callable.begin(), callable.end(),
                                  [](const Piece* p) { return p->type == CAT_REF && !p->keep_ref && p->allele_support == p->original_support; }),
                   callable.end());
    // ---- every callable allele: ProcessVariant, IsCallable && ShouldReport (:96-118)
    if (!callable.empty()) {
        std::vector<pb2_call_record> recs;
        std::vector<uint8_t> flags;
        rc = score_pieces(ctx, callable, arena, &recs, &flags, nullptr, nullptr);
        if (rc != PB2_OK) return rc;
        pc_.mark(6);
        for (size_t i = 0; i < recs.size(); i++) {
            if (!(flags[i] & (2 | 4))) continue;   // IsCallable && ShouldReport, or a forced allele (:108-118)
            called.push_back(recs[i]);
            pb2_call_record_ext e;
            memset(&e, 0, sizeof(e));
            for (int k = 0; k < 8; k++) e.collapsed_mut[k] = callable[i]->collapsed_mut[k];
            for (int k = 0; k < 3; k++) e.well_anchored_support[k] = callable[i]->wa[k];
            // An MNV candidate goes through ProcessVariant twice (AlleleCaller.cs:62-75, then :96-106 with every callable allele), and
            // CollapsedCoverageCalculator ADDS to ReadCollapsedCountTotal each time (CollapsedCoverageCalculator.cs:18-37): its totals come out doubled.
            // Marked here, applied where pb2_flush fills the totals in.
            if (callable[i]->from_candidate && callable[i]->type == CAT_MNV) e.collapsed_total[0] = kCollapsedTotalTwice;
            called_ext.push_back(e);
        }
    }
    pc_.mark(7);
    return PB2_OK;
}

// ------------------------------------------------------------------------------------------------ pb2_call_resident
namespace {
struct ResidentPlan {
    BatchCtx ctx;
    CandScoreArgs args;
    std::vector<uint8_t> arena;
    explicit ResidentPlan(pb2_handle* h) : ctx(h) {}
};
}  // namespace

void explicit_release_resident(pb2_handle* h) {
    if (h->resident_explicit) { delete static_cast<ResidentPlan*>(h->resident_explicit); h->resident_explicit = nullptr; }
}

bool explicit_resident_ready(pb2_handle* h) { return h->resident_explicit != nullptr; }

// First call: builds the plan (tables, gathers, scorer in append mode behind the hot kernel on the handle stream). Later calls (`side` given): replays the
// recorded launches on the side stream so that they run next to the hot kernel; the caller joins the streams and then prunes the reference records.
int explicit_call_resident(pb2_handle* h, Segment& seg, cudaStream_t side) {
    cudaStream_t st = h->stream;
    ResidentPlan* plan = static_cast<ResidentPlan*>(h->resident_explicit);
    if (plan == nullptr) {
        if (!h->forced.empty()) return pb2_fail(h, PB2_ERR_UNSUPPORTED, "pb2_call_resident: forced alleles are reported through pb2_flush");
        for (auto& c : h->cands) {
            if (!c.alive) continue;
            if (c.type == CAT_MNV || c.type == CAT_SNV) return pb2_fail(h, PB2_ERR_UNSUPPORTED, "pb2_call_resident: MNV/SNV candidates need the MNV reallocator; use pb2_flush");
            if (h->cfg.collapse && (c.open_left || c.open_right)) return pb2_fail(h, PB2_ERR_UNSUPPORTED, "pb2_call_resident: open-ended candidates need the collapser; use pb2_flush");
        }
        plan = new ResidentPlan(h);
        plan->ctx.use_resident_reads = true;
        h->resident_explicit = plan;
        std::deque<Piece> store;
        std::vector<Piece*> ps;
        std::vector<size_t> idx;
        for (size_t i = 0; i < h->cands.size(); i++) if (h->cands[i].alive) idx.push_back(i);
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return h->cands[a].position < h->cands[b].position; });
        for (size_t i : idx) { store.push_back(piece_of(h->cands[i])); ps.push_back(&store.back()); }
        h->arena.clear();
        const int rc = score_pieces(plan->ctx, ps, h->arena, nullptr, nullptr, nullptr, &seg);
        if (rc != PB2_OK) { explicit_release_resident(h); return rc; }
        memset(&plan->args, 0, sizeof(plan->args));
        CandScoreArgs& a = plan->args;
        a.cands = plan->ctx.d_cands.p; a.n = (int32_t)ps.size(); a.counts = plan->ctx.counts.p; a.collapsed = plan->ctx.collapsed.p;
        a.qsum = plan->ctx.want_q ? plan->ctx.qsum.p : nullptr; a.arena = plan->ctx.d_arena.p; a.chr_seq = h->d_chr; a.chr_len = h->chr_len;
        a.q_to_p_table = h->d_q_to_p; a.q_table_max = h->q_table_max; a.indel_repeat_filter = h->cfg.indel_repeat_filter;
        a.var_records = seg.var_records; a.var_count = seg.counters; a.var_capacity = seg.var_capacity;
        // replays run concurrently with the hot kernel: the scorer only flags what it called, explicit_prune_resident clears ref_valid afterwards
        a.ref_valid = nullptr;
        CUX(h, plan->ctx.d_flags.reserve(ps.size(), st, false, h));
        a.out_callable = plan->ctx.d_flags.p;
        return PB2_OK;
    }
    BatchCtx& ctx = plan->ctx;
    for (auto& g : ctx.gathers) {
        CUX(h, BatchCtx::gather_rows(h->segs[g.seg], ctx.req.p + g.req_off, g.n, ctx.counts.p + (size_t)g.row0 * kNumBins,
                                     ctx.collapsed.p + (size_t)g.row0 * kNumCollapsed, ctx.want_q ? ctx.qsum.p + (size_t)g.row0 * kNumBins : nullptr, h->dcfg.min_bq, side));
        h->total_launches += 1;
    }
    CUX(h, launch_score_candidates(plan->args, h->dcfg, side));
    h->total_launches += 1;
    return PB2_OK;
}
int explicit_prune_resident(pb2_handle* h, Segment& seg) {
    ResidentPlan* plan = static_cast<ResidentPlan*>(h->resident_explicit);
    if (plan == nullptr || seg.ref_valid == nullptr) return PB2_OK;
    CUX(h, launch_prune_ref_valid(plan->args.cands, plan->args.out_callable, plan->args.n, seg.ref_valid, h->stream));
    h->total_launches += 1;
    return PB2_OK;
}

// ------------------------------------------------------------------------------------------------ candidates of pushed reads
static int find_candidates_impl(pb2_handle* h, size_t first_read, const BatchHostView* host, int32_t snv_lo, int32_t snv_hi);
int explicit_find_candidates(pb2_handle* h, size_t first_read, const BatchHostView* host) { return find_candidates_impl(h, first_read, host, 0, 0); }
// CallMNVs off: SNV candidates are the counts — until RegionStateManager.AddCollapsableFromOtherBlocks (:441-457) pulls the finished SNV candidates of a
// later block into an earlier batch. What happens to them there depends on their open-end twins and the order they were raised in
// (RegionState.ExtractCollapsable :470-490 removes with List.Remove, i.e. the first candidate that Equals), so from then on the SNV candidates at the
// positions (lo, hi] are explicit: found in the kept reads exactly as the reference's finder raised them, and the hot kernel stops deriving SNVs from
// the counts there (Segment gapped/suppress array). No read that arrives later can touch these positions (reads come in position order).
int explicit_materialize_snvs(pb2_handle* h, int32_t lo, int32_t hi) {
    if (hi <= lo) return PB2_OK;
    // SNV candidates already in the table here can only be forced alleles; the reference adds a forced allele once calling has passed its position
    // (SmallVariantCaller.cs:99-104), i.e. after every read that covers it: keep that order in the position's candidate list
    std::vector<HostCand> forced_here;
    for (auto& c : h->cands)
        if (c.alive && c.type == CAT_SNV && c.position > lo && c.position <= hi) { forced_here.push_back(c); c.alive = false; }
    const int rc = find_candidates_impl(h, 0, nullptr, lo, hi);
    if (rc != PB2_OK) return rc;
    for (auto& c : forced_here) { c.alive = true; explicit_add_candidate(h, c); }
    h->snv_explicit_ranges.push_back({lo, hi});
    return PB2_OK;
}
// snv_lo < snv_hi: only the SNV candidates at positions in (snv_lo, snv_hi], found with the CallMNVs-off state machine (ShouldBuildUpMNV :170-181 returns
// false: every mismatch is its own SNV); used when count-based SNVs have to become explicit candidates (explicit_materialize_snvs)
static int find_candidates_impl(pb2_handle* h, size_t first_read, const BatchHostView* host, int32_t snv_lo, int32_t snv_hi) {
    const bool snv_only = snv_lo < snv_hi;
    const DeviceReads& R = h->reads;
    const size_t n = R.size() - first_read;
    if (n == 0) return PB2_OK;
    if (h->chr_len == 0) return PB2_OK;   // no reference: the finder has nothing to compare against (the reference would throw on refChromosome[...])
    cudaStream_t st = h->stream;
    CUX(h, cudaSetDevice(h->device));
    DevBuf<RawCand> d_raw; DevBuf<unsigned long long> d_count;
    // every insertion / deletion operation raises at most one candidate, every aligned base at most one SNV/MNV; the SNV/MNV bound is far from tight on
    // real data, so the buffer starts smaller and the kernel is run again with the exact size if it did not fit
    const bool per_base = h->cfg.call_mnvs || snv_only;
    const int64_t n_cig = R.n_cigar, n_seq = R.n_seq;
    // (a position-limited search - explicit_materialize_snvs - touches a few hundred positions)
    int64_t capacity = (snv_only ? 0 : n_cig) + (per_base ? std::min<int64_t>(n_seq, snv_only ? (1 << 18) : std::max<int64_t>(1 << 20, n_seq / 16)) : 0) + 16;
    CUX(h, d_count.reserve(1, st, false, h));
    const ReadsView rv = R.view();
    unsigned long long cnt = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        CUX(h, d_raw.reserve((size_t)capacity, st, false, h));
        CUX(h, cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), st));
        CUX(h, launch_reads_candidates(rv, (int32_t)first_read, h->d_chr, h->chr_len, h->dcfg.min_bq, snv_only ? 1 : h->cfg.call_mnvs, snv_only ? 0 : h->cfg.max_size_mnv,
                                       snv_only ? 0 : h->cfg.max_gap_mnv, h->cfg.expect_collapsed, d_raw.p, d_count.p, capacity, snv_only ? snv_lo : std::max(h->cleared_through, 0),
                                       snv_only ? snv_hi : INT32_MAX, snv_only ? 1 : 0, R.end_pos.p, st));
        h->total_launches += 1;
        CUX(h, cudaMemcpyAsync(&cnt, d_count.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        CUX(h, cudaStreamSynchronize(st));
        if ((int64_t)cnt <= capacity) break;
        if (attempt == 1) return pb2_fail(h, PB2_ERR_NOMEM, "candidate buffer overflow");
        capacity = (int64_t)cnt + 16;
    }
    if (cnt == 0) return PB2_OK;
    // Many occurrences, few distinct candidates: reduce by key on the device and bring back one row per distinct candidate (with the place of its first
    // occurrence). Small sets, and sets with alleles longer than RawCand::read_bases or a hash collision, take the one-by-one path below.
    std::vector<RawCand> raw;
    std::vector<CandGroup> groups;
    if (cnt >= 2048) {
        const int64_t n = (int64_t)cnt;
        DevBuf<unsigned long long> k0, k1; DevBuf<uint32_t> i0, i1; DevBuf<int32_t> head, gof, flg; DevBuf<uint8_t> tmp; DevBuf<CandGroup> dg;
        size_t tb = 0;
        CUX(h, cand_reduce_temp_bytes(n, &tb));
        CUX(h, k0.reserve((size_t)n, st, false, h)); CUX(h, k1.reserve((size_t)n, st, false, h)); CUX(h, i0.reserve((size_t)n, st, false, h)); CUX(h, i1.reserve((size_t)n, st, false, h));
        CUX(h, head.reserve((size_t)n, st, false, h)); CUX(h, gof.reserve((size_t)n, st, false, h)); CUX(h, flg.reserve(1, st, false, h)); CUX(h, tmp.reserve(tb + 16, st, false, h));
        CUX(h, cudaMemsetAsync(flg.p, 0, sizeof(int32_t), st));
        CUX(h, launch_cand_group(d_raw.p, n, k0.p, k1.p, i0.p, i1.p, head.p, gof.p, tmp.p, tb + 16, flg.p, st));
        int32_t n_groups = 0, flags = 0;
        CUX(h, cudaMemcpyAsync(&n_groups, gof.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CUX(h, cudaMemcpyAsync(&flags, flg.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CUX(h, cudaStreamSynchronize(st));
        h->total_launches += 5;
        if (flags == 0 && n_groups > 0) {
            CUX(h, dg.reserve((size_t)n_groups, st, false, h));
            CUX(h, launch_cand_reduce(d_raw.p, i1.p, gof.p, n, dg.p, n_groups, st));
            h->total_launches += 2;
            groups.resize((size_t)n_groups);
            CUX(h, cudaMemcpyAsync(groups.data(), dg.p, sizeof(CandGroup) * groups.size(), cudaMemcpyDeviceToHost, st));
            CUX(h, cudaStreamSynchronize(st));
            std::sort(groups.begin(), groups.end(), [](const CandGroup& a, const CandGroup& b) { return a.first_seen < b.first_seen; });
            // the head candidates themselves: gathered on the host side from the raw list (one small copy per group would be slower than the list)
            raw.resize((size_t)cnt);
            CUX(h, cudaMemcpy(raw.data(), d_raw.p, sizeof(RawCand) * (size_t)cnt, cudaMemcpyDeviceToHost));
            std::vector<RawCand> heads;
            heads.reserve(groups.size());
            for (auto& g : groups) heads.push_back(raw[g.first_index]);
            raw.swap(heads);
        }
    }
    if (groups.empty()) {
        raw.resize((size_t)cnt);
        CUX(h, cudaMemcpy(raw.data(), d_raw.p, sizeof(RawCand) * (size_t)cnt, cudaMemcpyDeviceToHost));
        // FindCandidates' own order: read by read, operation by operation
        std::sort(raw.begin(), raw.end(), [](const RawCand& a, const RawCand& b) { return a.read != b.read ? a.read < b.read : a.order < b.order; });
    }
    // alleles longer than RawCand::read_bases whose read is not in the caller's batch: fetched from the device store
    std::vector<uint8_t> fetched;
    auto read_bases_of = [&](const RawCand& rc, int n_from_read, std::string& out) -> int {
        if (n_from_read <= 8) { out.append(reinterpret_cast<const char*>(rc.read_bases), (size_t)n_from_read); return PB2_OK; }
        if (host != nullptr && (size_t)rc.read >= first_read) {
            const int64_t j = (int64_t)rc.read - (int64_t)first_read;
            out.append(reinterpret_cast<const char*>(host->bases) + (host->seq_off[j] - host->seq_lo) + rc.start_in_read, (size_t)n_from_read);
            return PB2_OK;
        }
        int64_t off = 0;
        CUX(h, cudaMemcpy(&off, R.seq_off.p + rc.read, sizeof(int64_t), cudaMemcpyDeviceToHost));
        fetched.resize((size_t)n_from_read);
        CUX(h, cudaMemcpy(fetched.data(), R.bases.p + off + rc.start_in_read, (size_t)n_from_read, cudaMemcpyDeviceToHost));
        out.append(reinterpret_cast<const char*>(fetched.data()), (size_t)n_from_read);
        return PB2_OK;
    };
    const char* chr = reinterpret_cast<const char*>(h->h_chr.data());
    for (size_t ci = 0; ci < raw.size(); ci++) {
        const RawCand& rc = raw[ci];
        if (rc.position <= h->cleared_through || rc.position < 1) continue;
        if (snv_only && (rc.type != CAT_SNV || rc.position <= snv_lo || rc.position > snv_hi)) continue;
        if ((int64_t)rc.position - 1 + rc.ref_len > h->chr_len) continue;   // Substring past the chromosome end throws in the reference
        HostCand c;
        c.position = rc.position; c.type = rc.type;
        c.open_left = (rc.flags & 1) != 0; c.open_right = (rc.flags & 2) != 0;
        c.ref.assign(chr + rc.position - 1, rc.ref_len);
        int rcode = PB2_OK;
        if (rc.type == CAT_INS) { c.alt.assign(1, chr[rc.position - 1]); rcode = read_bases_of(rc, (int)rc.alt_len - 1, c.alt); }
        else if (rc.type == CAT_DEL) c.alt.assign(1, chr[rc.position - 1]);
        else rcode = read_bases_of(rc, (int)rc.alt_len, c.alt);
        if (rcode != PB2_OK) return rcode;
        if (!groups.empty()) {   // one row per distinct candidate, its occurrences already summed
            const CandGroup& g = groups[ci];
            for (int k = 0; k < 3; k++) { c.support[k] = g.support[k]; c.well_anchored[k] = g.well_anchored[k]; }
            for (int k = 0; k < 8; k++) c.collapsed_mut[k] = g.collapsed_mut[k];
        } else {
            c.support[rc.dir] = 1;
            if (rc.flags & 4) c.well_anchored[rc.dir] = 1;
            if (rc.collapsed) {   // CandidateVariantFinder.Create (:352-384)
                const int t = rc.collapsed - 1;
                c.collapsed_mut[t]++;
                if (t == 4 || t == 6) c.collapsed_mut[2]++;
                else if (t == 5 || t == 7) c.collapsed_mut[3]++;
            }
        }
        explicit_add_candidate(h, c);
    }
    return PB2_OK;
}
