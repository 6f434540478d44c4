// Reads -> PVERT on the device (pb2_pvert.cuh): the staging side of pb2_push_reads / pb2_flush, and the 198-bin gather the explicit-candidate path
// and pb2_get_counts read from it.
//
// Follows (reference @ /root/reference):
//   RegionStateManager.AddAlleleCounts / GetAnchorType      src/lib/Pisces.Processing/RegionState/RegionStateManager.cs:83-220
//   Read.PositionMap / SequencedBaseDirectionMap             src/lib/Pisces.Domain/Models/Read.cs:390-421,535-562
//   CandidateVariantFinder.CheckDeletionQuality + the SNV open-end bookkeeping   src/lib/Pisces.Domain/Logic/CandidateVariantFinder.cs:90-203,294-320,496-553
//   CollapsedRegionState.AddCollapsedReadCount               src/lib/Pisces.Processing/RegionState/CollapsedRegionState.cs:28-44
#include <cub/device/device_scan.cuh>
#include "pb2_internal.hpp"

namespace pb2 {

namespace {
__device__ __forceinline__ bool pv_ref_span(int op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }   // M D N = X  (BamCommon.cs:560-573)
__device__ __forceinline__ bool pv_read_span(int op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }  // M I S = X  (:575-588)
__device__ __forceinline__ int pv_anchor_type(int end_pos, int base_pos, int start_pos) {   // RegionStateManager.GetAnchorType (:83-116), K = 5
    const int left = base_pos - start_pos, right = end_pos - base_pos;
    if (left >= right) return right >= kAnchorK ? kAnchorK : kNumAnchors - right - 1;
    return left >= kAnchorK ? kAnchorK : left;
}
__device__ __forceinline__ int64_t pv_locus_index(const RegionView& rg, int position) {
    if (position < rg.lo || position > rg.hi) return -1;
    return rg.index_of_pos ? (int64_t)rg.index_of_pos[position - rg.lo] : (int64_t)(position - rg.lo);
}
__device__ __forceinline__ int pv_allele2(uint8_t b) { return b == 'A' ? 0 : b == 'G' ? 1 : b == 'C' ? 2 : b == 'T' ? 3 : -1; }   // AlleleType order A G C T

// One row being written by a thread: the slots of the current (tile, class) stretch of its read, a 32-bit word (4 loci) at a time.
struct RowState {
    int64_t key;    // tile * n_classes + class, -1 = none
    int64_t row;
    int widx;       // word of the row being assembled, -1 = none
    int tag;        // deletion rows: the anchor bin all entries of the row share (row_meta holds one per row)
    uint32_t acc;
};

struct FillTargets {
    const int64_t* tile_row0;
    const int32_t* cls_end;
    int32_t* cursor;
    uint8_t* data;
    int2* row_meta;
    int32_t* row_amp;   // optional: amplicon id of the row's read (base rows)
    uint32_t* exc_entries;
    unsigned long long* exc_count;
    int64_t exc_capacity;
    const uint32_t* ref_slot_words;   // per locus one byte, allele2 << 6 of the reference base (1 where it is not A/C/G/T), four loci per word
};

// Walks one read exactly as RegionStateManager.AddAlleleCounts does, restricted to the reference positions [w_lo, w_hi] of ONE tile, and hands every entry
// there to the row writer: operations before the window are skipped whole, an aligned operation is walked over its bases inside the window only. One
// thread per (read, tile) piece: every thread does the same amount of work (at most 32 bases), whatever the phase of its read against the tile grid.
// kFill = false: only the rows are counted (same walk, same row boundaries: the two passes agree by construction).
template <bool kFill>
__device__ void pv_walk_piece(const ReadsView& rv, const RegionView& rg, int r, int end_pos, int w_lo, int w_hi, int n_classes, int32_t* cls_rows, const FillTargets& ft) {
    const int64_t c0 = rv.cigar_off[r], c1 = rv.cigar_off[r + 1];
    const int64_t s0 = rv.seq_off[r];
    const int read_len = (int)(rv.seq_off[r + 1] - s0);
    const int n_ops = (int)(c1 - c0);
    if (n_ops == 0) return;
    const int start_pos = rv.pos0[r] + 1;                                  // Read.Position; end_pos = Read.EndPosition (Read.cs:88-91)
    const bool reverse = (rv.flag[r] & 0x10) != 0;
    const int cg = (rg.expect_collapsed && rv.collapsed) ? pv_collapsed_group(rv.collapsed[r]) : 0;
    const uint8_t* __restrict__ bases = rv.bases + s0;
    const uint8_t* __restrict__ quals = rv.quals + s0;
    const uint8_t* __restrict__ dirs = rv.base_dirs ? rv.base_dirs + s0 : nullptr;
    auto dir_at = [&](int i) -> int { return dirs ? min((int)dirs[i], 2) : (reverse ? DIR_R : DIR_F); };
    auto del_q = [&](int idx) -> int {  // CandidateVariantFinder.CheckDeletionQuality (:294-320): min of the flanking qualities
        if (read_len == 0) return -1;
        const int after = idx < read_len ? quals[idx] : quals[idx - 1];
        const int before = idx > 0 ? quals[idx - 1] : after;
        return min(before, after);
    };
    // terminal deletion bookkeeping (:127-137)
    const int last_op = (int)(rv.cigar[c1 - 1] & 15);
    const int prev_op = n_ops >= 2 ? (int)(rv.cigar[c1 - 2] & 15) : -1;
    const bool ends_in_del = last_op == 2;
    const bool ends_in_del_before_clip = prev_op == 2 && last_op == 4;
    int del_len = 0, len_before_del = read_len;
    if (ends_in_del || ends_in_del_before_clip) {
        del_len = (int)(ends_in_del_before_clip ? (rv.cigar[c1 - 2] >> 4) : (rv.cigar[c1 - 1] >> 4));
        len_before_del = ends_in_del_before_clip ? read_len - (int)(rv.cigar[c1 - 1] >> 4) : read_len;
    }
    // open-end annotation (:496-553): first / last non-soft-clip operation, the last mapped position
    int first_op = (int)(rv.cigar[c0] & 15);
    if (first_op == 4 && n_ops >= 2) first_op = (int)(rv.cigar[c0 + 1] & 15);
    int last_nonclip = last_op;
    if (last_nonclip == 4 && n_ops >= 2) last_nonclip = prev_op;
    int max_mapped = -1;
    if (kFill) {
        int rp = start_pos;
        for (int i = 0; i < n_ops; i++) {
            const uint32_t c = rv.cigar[c0 + i];
            const int op = c & 15, len = (int)(c >> 4);
            if (pv_ref_span(op)) { if (pv_read_span(op) && len > 0) max_mapped = rp + len - 1; rp += len; }
        }
    }

    // two rows in flight, one per entry kind (separate variables: they stay in registers)
    RowState sb, sd;
    sb.key = sd.key = -1; sb.widx = sd.widx = -1; sb.acc = sd.acc = 0; sb.row = sd.row = 0; sb.tag = sd.tag = 0;
    auto flush_word = [&](RowState& s) {
        if (kFill && s.widx >= 0 && s.acc != 0) *reinterpret_cast<uint32_t*>(ft.data + s.row * 32 + s.widx * 4) |= s.acc;
        s.widx = -1; s.acc = 0;
    };
    // one entry into the row of its kind: `byte` is the slot value, meta what row_meta gets when the entry opens a row
    auto put_into = [&](RowState& s, int kind, int position, int dir, uint32_t byte, int2 meta, int tag) -> int64_t {
        const int64_t li = pv_locus_index(rg, position);
        if (li < 0) return -1;
        const int64_t tile = li >> 5;
        const int l = (int)(li & 31);
        const int cls = pv_class(kind, dir, cg);
        const int64_t key = tile * n_classes + cls;
        if (key != s.key || tag != s.tag) {
            flush_word(s);
            s.key = key;
            s.tag = tag;
            if (kFill) {
                const int k = atomicAdd(ft.cursor + key, 1);
                s.row = ft.tile_row0[tile] + (cls > 0 ? ft.cls_end[tile * n_classes + cls - 1] : 0) + k;
                ft.row_meta[s.row] = meta;
                if (ft.row_amp != nullptr && kind == 0) ft.row_amp[s.row] = rv.amplicon ? rv.amplicon[r] : -1;
            } else {
                atomicAdd(cls_rows + key, 1);
            }
        }
        if (kFill) {
            const int w = l >> 2;
            if (w != s.widx) { flush_word(s); s.widx = w; }
            s.acc |= byte << (8 * (l & 3));
        }
        return li;
    };
    // Deletion entries at positions [a, b] clipped to the window; a Deletion entry below the quality bar is never counted (:170-177)
    auto put_dels = [&](int a, int b, int dir, int dq, int anchor) {
        if (dq < rg.min_bq) return;
        for (int j = max(a, w_lo); j <= min(b, w_hi); j++) put_into(sd, 1, j, dir, (uint32_t)min(max(dq, 1), 63), make_int2(INT32_MIN + anchor, 0), anchor);
    };
    const int2 read_meta = make_int2(start_pos, end_pos);

    int read_idx = 0, ref_pos = start_pos, last_position = start_pos - 1;
    for (int oi = 0; oi < n_ops; oi++) {
        const uint32_t c = rv.cigar[c0 + oi];
        const int op = c & 15, len = (int)(c >> 4);
        const bool rs = pv_read_span(op), fs = pv_ref_span(op);
        if (rs && !fs) {                                                          // I / S: not mapped to the reference
            if (ends_in_del_before_clip && read_idx <= len_before_del && len_before_del < read_idx + len)   // (:148-159): at the first clipped base
                put_dels(last_position + 1, last_position + del_len, dir_at(len_before_del), del_q(len_before_del), kNumAnchors - 1);
            read_idx += len;
        } else if (rs && fs) {
            if (len > 0) {
                if (ref_pos > last_position + 1 && ref_pos - 1 >= w_lo && last_position + 1 <= w_hi)       // deletion (or N skip) before this base (:170-177)
                    put_dels(last_position + 1, ref_pos - 1, dir_at(read_idx), del_q(read_idx), pv_anchor_type(end_pos, ref_pos, start_pos));
                const int k0 = max(0, w_lo - ref_pos), k1 = min(len, w_hi - ref_pos + 1);
                // the flagged-entry test of one counted base (fill pass): SNV-candidate bookkeeping the counts cannot express (CallMNVs off;
                // CandidateVariantFinder.cs:90-168). Only a usable mismatch against an A/C/G/T reference base can matter; it goes to the segment's side list
                // (the rule of tile_scatter_kernel)
                auto flag_entry = [&](int k, int ri, int position, int dir, int a2, int q, int64_t li) {
                    if (a2 < 0 || q < rg.min_bq) return;
                    const bool in_chr = rg.chr == nullptr || position <= rg.chr_len;
                    const uint8_t rb = (rg.chr != nullptr && position >= 1 && position <= rg.chr_len) ? rg.chr[position - 1] : (uint8_t)'N';
                    const int ra2 = pv_allele2(rb);
                    if (ra2 < 0 || ra2 == a2) return;
                    uint32_t code = 0;
                    if (op != 0 || !in_chr) code |= PB2_ENTRY_NO_CANDIDATE;
                    else {
                        if (k + 1 < len) {   // open on the right: the next base of this operation exists and is unusable -> FlushVariant(..., openRight = true)
                            const int nq = quals[ri + 1];
                            const bool nb_n = pv_allele2(bases[ri + 1]) < 0;
                            bool nref_n = false, n_in = true;
                            if (rg.chr) {
                                n_in = position + 1 <= rg.chr_len;
                                if (n_in) nref_n = pv_allele2(rg.chr[position]) < 0;
                            }
                            if (n_in && (nq < rg.min_bq || nb_n || nref_n)) code |= PB2_ENTRY_OPEN_RIGHT;
                        }
                        if (first_op == 0 && position == start_pos) code |= PB2_ENTRY_OPEN_LEFT;
                        if (last_nonclip == 0 && position == max_mapped) code |= PB2_ENTRY_OPEN_RIGHT;
                    }
                    if (code) {
                        const unsigned long long slot = atomicAdd(ft.exc_count, 1ull);
                        if ((int64_t)slot < ft.exc_capacity) {
                            const int an = pv_anchor_type(end_pos, position, start_pos);
                            const int cc = pv_collapsed_code(cg, dir);
                            ft.exc_entries[2 * slot] = (uint32_t)li;
                            ft.exc_entries[2 * slot + 1] = (code | (uint32_t)a2 | ((uint32_t)dir << 3)) | ((uint32_t)min(q, 127) << 8) | ((uint32_t)(an | (cc << 4)) << 16);
                        }
                    }
                };
                // the common case: every position is a locus (no interval gaps) and the bases of the stretch share one direction. The stretch is one row
                // of its tile: the row is opened once and written word by word - all threads of a warp run the same eight iterations over the tile's words
                bool fast = k0 < k1 && rg.index_of_pos == nullptr;
                const int dir0 = fast ? dir_at(read_idx + k0) : 0;
                if (fast && dirs != nullptr) for (int k = k0 + 1; k < k1; k++) fast = fast && dir_at(read_idx + k) == dir0;
                if (fast) {
                    const int64_t tile = (int64_t)(w_lo - rg.lo) >> 5;
                    const int cls = pv_class(0, dir0, cg);
                    const int64_t key = tile * n_classes + cls;
                    if (key != sb.key || sb.tag != 0) {
                        flush_word(sb);
                        sb.key = key; sb.tag = 0;
                        if (kFill) {
                            const int kk = atomicAdd(ft.cursor + key, 1);
                            sb.row = ft.tile_row0[tile] + (cls > 0 ? ft.cls_end[tile * n_classes + cls - 1] : 0) + kk;
                            ft.row_meta[sb.row] = read_meta;
                            if (ft.row_amp != nullptr) ft.row_amp[sb.row] = rv.amplicon ? rv.amplicon[r] : -1;
                        } else {
                            atomicAdd(cls_rows + key, 1);
                        }
                    }
                    if (kFill) {
                        flush_word(sb);
                        const int tp = rg.lo + (int)(tile << 5);                       // position of the tile's locus 0
                        const int la = ref_pos + k0 - tp, lb = ref_pos + k1 - 1 - tp;   // loci of the stretch inside the tile
                        const int delta = read_idx - ref_pos + tp;                      // read index of the base at locus l: l + delta
                        uint32_t* const rowp = reinterpret_cast<uint32_t*>(ft.data + sb.row * 32);
                        const uint8_t* const sl = rv.slots + s0;                        // the read's slot bytes (16 bytes of slack around the plane)
                        const uint32_t* const refw = ft.ref_slot_words + (tile << 3);
#pragma unroll 1
                        for (int w = 0; w < 8; w++) {
                            const int l4 = 4 * w;
                            if (l4 + 3 < la || l4 > lb) continue;
                            // four slot bytes from an arbitrary byte address: two aligned words and a funnel shift
                            const uintptr_t a = reinterpret_cast<uintptr_t>(sl + (l4 + delta));
                            const uint32_t* const aw = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                            uint32_t v = __funnelshift_r(aw[0], aw[1], (unsigned)(a & 3u) * 8u);
                            const int c0 = max(la - l4, 0), c1 = min(lb - l4, 3);       // bytes of the word inside the stretch
                            const uint32_t m = (0xffffffffu << (8 * c0)) & (0xffffffffu >> (8 * (3 - c1)));
                            v &= m;
                            // candidate flags: only a base whose allele differs from the reference allele can need one
                            uint32_t x = rg.chr != nullptr ? ((v ^ refw[w]) & 0xc0c0c0c0u & m) : 0u;
                            while (x) {
                                const int j = (__ffs((int)x) - 1) >> 3;
                                x &= ~(0xffu << (8 * j));
                                const int l = l4 + j, position = tp + l, k = position - ref_pos, ri = read_idx + k;
                                flag_entry(k, ri, position, dir0, pv_allele2(bases[ri]), quals[ri], (int64_t)(position - rg.lo));
                            }
                            if (c0 == 0 && c1 == 3) rowp[w] = v;   // a whole word of this stretch
                            else if (v) rowp[w] |= v;              // the row is this thread's own and starts out zero: a later stretch may complete the word
                        }
                    }
                } else
                for (int k = k0; k < k1; k++) {
                    const int ri = read_idx + k, position = ref_pos + k;
                    const int dir = dir_at(ri);
                    const uint8_t b = bases[ri];
                    const int q = quals[ri];
                    const int a2 = pv_allele2(b);
                    const uint32_t byte = a2 < 0 ? 1u : ((uint32_t)a2 << 6) | (uint32_t)min(max(q, 1), 63);
                    const int64_t li = put_into(sb, 0, position, dir, byte, read_meta, 0);
                    if (!kFill || li < 0) continue;
                    flag_entry(k, ri, position, dir, a2, q, li);
                }
                last_position = ref_pos + len - 1;
            }
            read_idx += len;
            ref_pos += len;
            if (ref_pos > w_hi + 1 && !ends_in_del && !ends_in_del_before_clip) break;   // everything further lies behind the window
        } else if (fs) {
            ref_pos += len;
        }
    }
    if (ends_in_del && read_len > 0)                                              // (:195-210)
        put_dels(last_position + 1, last_position + del_len, dir_at(read_len - 1), del_q(read_len - 1), kNumAnchors - 1);
    flush_word(sb);
    flush_word(sd);
}

// One thread per (read, tile) piece: thread t takes read t / kWalkPieces and, of the tiles the read touches, the (t % kWalkPieces)-th, (+ kWalkPieces)-th, ...
constexpr int kWalkPieces = 8;
// A SIMPLE read: one 'M' operation, its own direction for every base, every position a locus. Its pieces are written by pvert_fill_simple_kernel.
__device__ __forceinline__ bool pv_simple_read(const ReadsView& rv, const RegionView& rg, int r) {
    if (rv.base_dirs != nullptr || rg.index_of_pos != nullptr) return false;
    const int64_t c0 = rv.cigar_off[r];
    return rv.cigar_off[r + 1] - c0 == 1 && (rv.cigar[c0] & 15u) == 0;
}
// complex: the reads that are not simple, listed by pvert_count_simple_kernel (complex[0] = how many, then their indices)
template <bool kFill>
__global__ void __launch_bounds__(256) pvert_walk_kernel(ReadsView rv, RegionView rg, const int32_t* __restrict__ end_pos_of, int n_classes, int32_t* __restrict__ cls_rows,
                                                        FillTargets ft, const int32_t* __restrict__ complex) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = (int)(t % kWalkPieces);
    if (t / kWalkPieces >= complex[0]) return;
    const int r = complex[1 + t / kWalkPieces];
    const int start_pos = rv.pos0[r] + 1, end_pos = end_pos_of[r];
    const int a = max(start_pos, rg.lo), e = min(end_pos, rg.hi);
    if (a > e) return;
    // loci of the staged window the read touches: [l0, l1]
    const int64_t l0 = rg.index_ge ? rg.index_ge[a - rg.lo] : (int64_t)(a - rg.lo);
    const int64_t l1 = (rg.index_ge ? (int64_t)rg.index_ge[e - rg.lo + 1] : (int64_t)(e - rg.lo + 1)) - 1;
    if (l0 > l1) return;
    for (int64_t tile = (l0 >> 5) + k; tile <= (l1 >> 5); tile += kWalkPieces) {
        const int64_t f = tile << 5, g = min(f + 31, rg.n_loci - 1);
        const int w_lo = rg.positions ? rg.positions[f] : rg.lo + (int)f, w_hi = rg.positions ? rg.positions[g] : rg.lo + (int)g;
        pv_walk_piece<kFill>(rv, rg, r, end_pos, max(w_lo, a), min(w_hi, e), n_classes, cls_rows, ft);
    }
}
// The candidate-flag test of one counted base of a simple read (see flag_entry in pv_walk_piece), out of line: it runs for the few bases whose allele
// differs from the reference allele.
__device__ __noinline__ void pv_flag_simple(const uint8_t* __restrict__ rbases, const uint8_t* __restrict__ rquals, const uint8_t* __restrict__ chr, int64_t chr_len, int min_bq,
                                            int region_lo, uint32_t* __restrict__ exc_entries, unsigned long long* __restrict__ exc_count, int64_t exc_capacity, int ri,
                                            int position, int dir, int cg, int start_pos, int end_pos, int len) {
    const int a2 = pv_allele2(rbases[ri]);
    const int q = rquals[ri];
    if (a2 < 0 || q < min_bq) return;
    const bool in_chr = position <= chr_len;
    const uint8_t rb = (position >= 1 && position <= chr_len) ? chr[position - 1] : (uint8_t)'N';
    const int ra2 = pv_allele2(rb);
    if (ra2 < 0 || ra2 == a2) return;
    uint32_t code = 0;
    if (!in_chr) code |= PB2_ENTRY_NO_CANDIDATE;
    else {
        if (ri + 1 < len) {   // open on the right: the next base of the operation exists and is unusable -> FlushVariant(..., openRight = true)
            const int nq = rquals[ri + 1];
            const bool nb_n = pv_allele2(rbases[ri + 1]) < 0;
            const bool n_in = position + 1 <= chr_len;
            const bool nref_n = n_in && pv_allele2(chr[position]) < 0;
            if (n_in && (nq < min_bq || nb_n || nref_n)) code |= PB2_ENTRY_OPEN_RIGHT;
        }
        if (position == start_pos) code |= PB2_ENTRY_OPEN_LEFT;
        if (position == end_pos) code |= PB2_ENTRY_OPEN_RIGHT;
    }
    if (!code) return;
    const unsigned long long slot = atomicAdd(exc_count, 1ull);
    if ((int64_t)slot < exc_capacity) {
        const int an = pv_anchor_type(end_pos, position, start_pos);
        const int cc = pv_collapsed_code(cg, dir);
        exc_entries[2 * slot] = (uint32_t)(position - region_lo);
        exc_entries[2 * slot + 1] = (code | (uint32_t)a2 | ((uint32_t)dir << 3)) | ((uint32_t)min(q, 127) << 8) | ((uint32_t)(an | (cc << 4)) << 16);
    }
}

// The pieces of simple reads (the bulk of any read set): a piece is one row, copied out of the read's slot bytes a word at a time - the nine source
// words are requested together, so a thread waits for memory once per piece. Few registers: many pieces in flight per SM.
// Thread mapping: blockIdx.x * 256 + threadIdx.x = which read, its pieces k = blockIdx.y, + gridDim.y, ... in turn (launched with gridDim.y = 1: every piece of
// a read by one thread): the 32 reads of a warp are neighbours in position order, their k-th pieces fall into one or two tiles, and the warp takes its rows from a tile's cursor with ONE atomic per (tile, class) instead of 32 (the
// reads being sorted, every tile's cursor is hammered by the few thousand threads in flight around it: the same-address atomics were the kernel's time).
__global__ void __launch_bounds__(256) pvert_fill_simple_kernel(ReadsView rv, RegionView rg, int n_classes, FillTargets ft) {
    const int r = blockIdx.x * 256 + threadIdx.x, k = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const bool mine = r < rv.n_reads && pv_simple_read(rv, rg, r);
    int len = 0, start_pos = 0, end_pos = 0, a = 1, e = 0, dir = 0, cg = 0;
    int64_t s0 = 0;
    if (mine) {
        len = (int)(rv.cigar[rv.cigar_off[r]] >> 4);
        start_pos = rv.pos0[r] + 1; end_pos = rv.pos0[r] + len;
        a = max(start_pos, rg.lo); e = min(end_pos, rg.hi);
        if (len == 0) e = a - 1;
        s0 = rv.seq_off[r];
        dir = (rv.flag[r] & 0x10) ? DIR_R : DIR_F;
        cg = (rg.expect_collapsed && rv.collapsed) ? pv_collapsed_group(rv.collapsed[r]) : 0;
    }
    const int cls = pv_class(0, dir, cg);
    int tile = mine && a <= e ? ((a - rg.lo) >> 5) + k : 1;
    const int tile_last = mine && a <= e ? ((e - rg.lo) >> 5) : 0;
    while (__any_sync(0xffffffffu, tile <= tile_last)) {
        const bool work = tile <= tile_last;
        const long long key = work ? (long long)tile * n_classes + cls : -1 - lane;   // idle lanes: keys of their own
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        int kk = 0;
        if (work) {
            const int leader = __ffs((int)peers) - 1;
            if (lane == leader) kk = atomicAdd(ft.cursor + key, __popc(peers));
            kk = __shfl_sync(peers, kk, leader) + __popc(peers & ((1u << lane) - 1u));
            const int tp = rg.lo + (tile << 5);
            const int la = max(a - tp, 0), lb = min(e - tp, 31);
            const int delta = tp - start_pos;                                         // read index of the base at locus l: l + delta
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(rv.slots + s0 + delta);   // address of the slot byte of locus 0 (may lie before the read: never loaded)
            const uint32_t* const aw = reinterpret_cast<const uint32_t*>(a0 & ~(uintptr_t)3);
            const unsigned sh = (unsigned)(a0 & 3u) * 8u;
            const int wa = la >> 2, wb = lb >> 2;
            uint32_t src[9];
#pragma unroll
            for (int w = 0; w < 9; w++) src[w] = (w >= wa && w <= wb + 1) ? aw[w] : 0u;
            const int64_t row = ft.tile_row0[tile] + (cls > 0 ? ft.cls_end[(int64_t)tile * n_classes + cls - 1] : 0) + kk;
            ft.row_meta[row] = make_int2(start_pos, end_pos);
            if (ft.row_amp != nullptr) ft.row_amp[row] = rv.amplicon ? rv.amplicon[r] : -1;
            const uint32_t* const refw = ft.ref_slot_words + ((int64_t)tile << 3);
            uint32_t out[8];
#pragma unroll
            for (int w = 0; w < 8; w++) {
                out[w] = 0;
                if (w < wa || w > wb) continue;
                const int c0 = max(la - 4 * w, 0), c1 = min(lb - 4 * w, 3);            // bytes of the word inside the read
                const uint32_t m = (0xffffffffu << (8 * c0)) & (0xffffffffu >> (8 * (3 - c1)));
                const uint32_t v = __funnelshift_r(src[w], src[w + 1], sh) & m;
                uint32_t x = rg.chr != nullptr ? ((v ^ refw[w]) & 0xc0c0c0c0u & m) : 0u;  // only a base whose allele differs from the reference allele can need a flag
                while (x) {
                    const int j = (__ffs((int)x) - 1) >> 3;
                    x &= ~(0xffu << (8 * j));
                    const int l = 4 * w + j;
                    pv_flag_simple(rv.bases + s0, rv.quals + s0, rg.chr, rg.chr_len, rg.min_bq, rg.lo, ft.exc_entries, ft.exc_count, ft.exc_capacity, l + delta, tp + l, dir, cg,
                                   start_pos, end_pos, len);
                }
                out[w] = v;
            }
            // the whole row in two 16-byte stores (empty slots are zero: whole sectors are written, nothing is read back)
            uint4* const rowq = reinterpret_cast<uint4*>(ft.data + row * 32);
            rowq[0] = make_uint4(out[0], out[1], out[2], out[3]);
            rowq[1] = make_uint4(out[4], out[5], out[6], out[7]);
        }
        tile += (int)gridDim.y;
    }
}

// the rows simple reads will take, same mapping and the same one-atomic-per-group trick
__global__ void __launch_bounds__(256) pvert_count_simple_kernel(ReadsView rv, RegionView rg, int n_classes, int32_t* __restrict__ cls_rows, int32_t* __restrict__ complex) {
    const int r = blockIdx.x * 256 + threadIdx.x, k = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const bool mine = r < rv.n_reads && pv_simple_read(rv, rg, r);
    if (k == 0) {   // the other reads go to the list the general walker works from (one atomic per warp)
        const unsigned cx = __ballot_sync(0xffffffffu, r < rv.n_reads && !mine);
        if (cx) {
            int base = 0;
            if (lane == __ffs((int)cx) - 1) base = atomicAdd(complex, __popc(cx));
            base = __shfl_sync(0xffffffffu, base, __ffs((int)cx) - 1);
            if ((cx >> lane) & 1u) complex[1 + base + __popc(cx & ((1u << lane) - 1u))] = r;
        }
    }
    int a = 1, e = 0, cls = 0;
    if (mine) {
        const int len = (int)(rv.cigar[rv.cigar_off[r]] >> 4);
        a = max(rv.pos0[r] + 1, rg.lo); e = min(rv.pos0[r] + len, rg.hi);
        if (len == 0) e = a - 1;
        const int cg = (rg.expect_collapsed && rv.collapsed) ? pv_collapsed_group(rv.collapsed[r]) : 0;
        cls = pv_class(0, (rv.flag[r] & 0x10) ? DIR_R : DIR_F, cg);
    }
    int tile = mine && a <= e ? ((a - rg.lo) >> 5) + k : 1;
    const int tile_last = mine && a <= e ? ((e - rg.lo) >> 5) : 0;
    while (__any_sync(0xffffffffu, tile <= tile_last)) {
        const bool work = tile <= tile_last;
        const long long key = work ? (long long)tile * n_classes + cls : -1 - lane;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (work && lane == __ffs((int)peers) - 1) atomicAdd(cls_rows + key, __popc(peers));
        tile += (int)gridDim.y;
    }
}

// complex: [1 + n_reads] ints, complex[0] zeroed by the caller before the count pass; n_complex: how many the count pass listed (fill pass), or -1: not
// known on the host yet (count pass: the kernel is launched for every read and the surplus threads leave at once)
template <bool kFill>
static cudaError_t launch_walk(const ReadsView& rv, const RegionView& rg, const int32_t* end_pos, int n_classes, int32_t* cls_rows, const FillTargets& ft, int32_t* complex,
                               int64_t n_complex, cudaStream_t st) {
    // One thread walks ALL pieces of its read (grid.y = 1), one after the other: the lanes of a warp are then still at their k-th pieces together (one
    // aggregated atomic per tile and class), the read's position / CIGAR / offsets are read once, and the 36-byte windows of consecutive pieces overlap in
    // L1. Measured against a piece per block row (grid.y = 8, 4, 2): 1.47 / 1.20 / 0.98 -> 0.83 ms, 4.14 -> 1.18 GB read from DRAM for 0.65 GB of input.
    const dim3 sgrid((unsigned)((rv.n_reads + 255) / 256), 1);
    if (kFill) pvert_fill_simple_kernel<<<sgrid, 256, 0, st>>>(rv, rg, n_classes, ft);
    else pvert_count_simple_kernel<<<sgrid, 256, 0, st>>>(rv, rg, n_classes, cls_rows, complex);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int64_t threads = (n_complex < 0 ? (int64_t)rv.n_reads : n_complex) * kWalkPieces;
    if (threads > 0) pvert_walk_kernel<kFill><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(rv, rg, end_pos, n_classes, cls_rows, ft, complex);
    return cudaGetLastError();
}

// rows per class -> inclusive prefix of the padded (multiple of 32) rows inside the tile; tile_rows[t] = rows of the tile
__global__ void pvert_layout_kernel(int32_t* __restrict__ cls_rows, int32_t n_tiles, int n_classes, int64_t* __restrict__ tile_rows) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == n_tiles) { tile_rows[t] = 0; return; }
    int32_t acc = 0;
    for (int c = 0; c < n_classes; c++) {
        acc += (cls_rows[(int64_t)t * n_classes + c] + 31) & ~31;
        cls_rows[(int64_t)t * n_classes + c] = acc;
    }
    tile_rows[t] = acc;
}

// One warp per 1 KB block, in place: 32 rows x 32 slot bytes -> per lane (= locus) the eight bit planes of its 32 slots.
constexpr int kTransposeWarps = 8;
__global__ void __launch_bounds__(32 * kTransposeWarps) pvert_transpose_kernel(uint8_t* __restrict__ data, int64_t n_blocks) {
    __shared__ __align__(16) uint8_t s_col[kTransposeWarps][32][36];   // [locus][row], padded: 9 words per locus, conflict-free column reads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * kTransposeWarps + warp;
    if (b >= n_blocks) return;
    uint8_t* blk = data + b * 1024;
    // lane = row: its 32 slot bytes
    const uint4 r0 = *reinterpret_cast<const uint4*>(blk + lane * 32), r1 = *reinterpret_cast<const uint4*>(blk + lane * 32 + 16);
    const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    if (__all_sync(0xffffffffu, (rw[0] | rw[1] | rw[2] | rw[3] | rw[4] | rw[5] | rw[6] | rw[7]) == 0)) return;   // an empty block stays all zero
#pragma unroll
    for (int l = 0; l < 32; l++) s_col[warp][l][lane] = (uint8_t)(rw[l >> 2] >> (8 * (l & 3)));
    __syncwarp();
    // lane = locus: its column, rows 0..31, eight rows per 64-bit word
    uint32_t plane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const uint32_t lo = *reinterpret_cast<const uint32_t*>(&s_col[warp][lane][8 * m]), hi = *reinterpret_cast<const uint32_t*>(&s_col[warp][lane][8 * m + 4]);
        const unsigned long long t = pv_transpose8x8((unsigned long long)lo | ((unsigned long long)hi << 32));
#pragma unroll
        for (int k = 0; k < 8; k++) plane[k] |= (uint32_t)((t >> (8 * k)) & 0xffull) << (8 * m);   // bit k of rows 8m..8m+7
    }
    __syncwarp();
    // slot byte: bit 7 b1, bit 6 b0, bits 5..0 quality.  half 0: B0 B1 Q5 Q4, half 1: Q3 Q2 Q1 Q0
    *reinterpret_cast<uint4*>(blk + lane * 16) = make_uint4(plane[6], plane[7], plane[5], plane[4]);
    *reinterpret_cast<uint4*>(blk + 512 + lane * 16) = make_uint4(plane[3], plane[2], plane[1], plane[0]);
}

__global__ void pvert_ref_bases_kernel(const uint8_t* __restrict__ chr, int64_t chr_len, const int32_t* __restrict__ positions, int32_t first_position, int64_t n_loci,
                                       uint8_t* __restrict__ ref_base, uint8_t* __restrict__ ref_slot, int64_t n_padded) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_padded) return;
    uint8_t b = 'N';
    if (i < n_loci) {
        const int64_t p = positions ? positions[i] : (int64_t)first_position + i;
        if (chr != nullptr && p >= 1 && p <= chr_len) b = chr[p - 1];
        ref_base[i] = b;
    }
    if (ref_slot != nullptr) { const int a2 = pv_allele2(b); ref_slot[i] = a2 < 0 ? (uint8_t)1 : (uint8_t)(a2 << 6); }
}

// One warp per requested locus: lane j takes row 32 g + j of every 32-row group of every class of the locus' tile.
constexpr int kPvGatherWarps = 4;
__global__ void __launch_bounds__(32 * kPvGatherWarps)
pvert_gather_kernel(PvertPileup in, const int32_t* __restrict__ req_locus, int32_t n_req, int32_t* __restrict__ out_counts, int32_t* __restrict__ out_collapsed, int min_bq) {
    __shared__ int s_hist[kPvGatherWarps][kNumBins + kNumCollapsed];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * kPvGatherWarps + warp;
    if (r >= n_req) return;
    int* hist = s_hist[warp];
    for (int b = lane; b < kNumBins + kNumCollapsed; b += 32) hist[b] = 0;
    __syncwarp();
    const int64_t locus = req_locus[r];
    if (locus >= 0 && locus < in.n_loci) {
        const int64_t tile = locus >> 5;
        const int l = (int)(locus & 31);
        const int32_t position = in.positions ? in.positions[locus] : in.first_position + (int32_t)locus;
        const int64_t row0 = in.tile_row0[tile];
        int32_t prev_end = 0;
        for (int c = 0; c < in.n_classes; c++) {
            const int32_t end = in.cls_end[tile * in.n_classes + c];
            const int kind = pv_class_kind(c), dir = pv_class_dir(c), ct = pv_collapsed_code(pv_class_cg(c), dir);
            for (int32_t g = prev_end; g < end; g += 32) {
                const uint8_t* blk = in.data + (row0 + g) * 32;
                const uint4 x = *reinterpret_cast<const uint4*>(blk + l * 16), y = *reinterpret_cast<const uint4*>(blk + 512 + l * 16);
                const uint32_t b0 = (x.x >> lane) & 1u, b1 = (x.y >> lane) & 1u;
                const uint32_t q = (((x.z >> lane) & 1u) << 5) | (((x.w >> lane) & 1u) << 4) | (((y.x >> lane) & 1u) << 3) | (((y.y >> lane) & 1u) << 2) |
                                   (((y.z >> lane) & 1u) << 1) | ((y.w >> lane) & 1u);
                if ((b0 | b1 | q) == 0) continue;   // empty slot
                const int2 meta = in.row_meta[row0 + g + lane];
                int allele, an;
                if (kind == 1) { allele = AT_DEL; an = meta.x - INT32_MIN; }
                else { allele = (int)q < min_bq ? AT_N : (int)(b0 | (b1 << 1)); an = pv_anchor_type(meta.y, position, meta.x); }   // RegionStateManager.cs:180-181
                atomicAdd(&hist[(allele * kNumDirs + dir) * kNumAnchors + an], 1);
                if (allele != AT_N && ct != 0) {   // CollapsedRegionState.AddCollapsedReadCount (:28-44)
                    atomicAdd(&hist[kNumBins + ct - 1], 1);
                    if (ct - 1 == 4 || ct - 1 == 6) atomicAdd(&hist[kNumBins + 2], 1);
                    else if (ct - 1 == 5 || ct - 1 == 7) atomicAdd(&hist[kNumBins + 3], 1);
                }
            }
            prev_end = end;
        }
        __syncwarp();
    }
    for (int b = lane; b < kNumBins; b += 32) out_counts[(int64_t)r * kNumBins + b] = hist[b];
    if (out_collapsed != nullptr && lane < kNumCollapsed) out_collapsed[(int64_t)r * kNumCollapsed + lane] = hist[kNumBins + lane];
}

// AmpliconBiasCalculator (src/lib/Pisces.Calculators/AmpliconBiasCalculator.cs:20-133) for the SNV records of a record stream. One warp per record: it
// walks the base rows of the record's tile exactly as the gather above, tallying per amplicon name the usable bases of the locus (CoverageByAmplicon:
// RegionState.AddAmpliconCount from RegionStateManager.cs:188, every base counted as A/C/G/T) and those equal to the alternate base (SupportByAmplicon:
// the SNV candidates of CandidateVariantFinder.cs:205-232 merged by RegionState.AddCandidate :138-170 - for count-based SNVs the same reads), in at
// most Constants.MaxNumOverlappingAmplicons = 6 slots. The verdict BiasDetected is an OR over the amplicons, so the reference's first-seen slot order
// does not matter; a seventh name at a called SNV is the reference's IndexOutOfRangeException (status 1).
constexpr int kAmpSlots = 6;
constexpr int kAmpWarps = 4;
__global__ void __launch_bounds__(32 * kAmpWarps)
pvert_amplicon_bias_kernel(PvertPileup in, const int32_t* __restrict__ row_amp, pb2_call_record* __restrict__ records, const unsigned long long* __restrict__ n_dev, int64_t n_host,
                           int64_t capacity, const uint8_t* __restrict__ valid, float acceptance, int min_bq, int* __restrict__ status) {
    __shared__ int s_name[kAmpWarps][kAmpSlots], s_cov[kAmpWarps][kAmpSlots], s_sup[kAmpWarps][kAmpSlots], s_n[kAmpWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t n = n_dev ? (int64_t)*n_dev : n_host;
    if (n > capacity) n = capacity;
    for (int64_t i = (int64_t)blockIdx.x * kAmpWarps + warp; i < n; i += (int64_t)gridDim.x * kAmpWarps) {
        const pb2_call_record& rec = records[i];
        if (valid != nullptr && !(valid[i] & 1)) continue;
        if (rec.type != CAT_SNV || rec.allele_support <= 0 || rec.ref_len != 1 || rec.alt_len != 1) continue;   // Compute :20-31 (AlleleCaller.cs:217-228)
        const int alt2 = pv_allele2((uint8_t)(rec.allele_bytes >> 8));
        int64_t locus = -1;
        if (in.positions == nullptr) locus = (int64_t)rec.position - in.first_position;
        else {
            int64_t lo = 0, hi = in.n_loci;
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (in.positions[mid] < rec.position) lo = mid + 1; else hi = mid; }
            if (lo < in.n_loci && in.positions[lo] == rec.position) locus = lo;
        }
        if (locus < 0 || locus >= in.n_loci || alt2 < 0) continue;
        if (lane == 0) s_n[warp] = 0;
        __syncwarp();
        const int64_t tile = locus >> 5;
        const int l = (int)(locus & 31);
        const int64_t row0 = in.tile_row0[tile];
        int32_t prev_end = 0;
        bool overflow = false;
        for (int c = 0; c < in.n_classes; c++) {
            const int32_t end = in.cls_end[tile * in.n_classes + c];
            if (pv_class_kind(c) == 0) {
                for (int32_t g = prev_end; g < end; g += 32) {
                    const uint8_t* blk = in.data + (row0 + g) * 32;
                    const uint4 x = *reinterpret_cast<const uint4*>(blk + l * 16), y = *reinterpret_cast<const uint4*>(blk + 512 + l * 16);
                    const uint32_t b0 = (x.x >> lane) & 1u, b1 = (x.y >> lane) & 1u;
                    const uint32_t q = (((x.z >> lane) & 1u) << 5) | (((x.w >> lane) & 1u) << 4) | (((y.x >> lane) & 1u) << 3) | (((y.y >> lane) & 1u) << 2) |
                                       (((y.z >> lane) & 1u) << 1) | ((y.w >> lane) & 1u);
                    const int amp = row_amp[row0 + g + lane];
                    const bool usable = (int)q >= min_bq && amp >= 0;      // an N base is (0, q = 1) and min_bq >= 2; empty slots have q = 0
                    const bool is_alt = usable && (int)(b0 | (b1 << 1)) == alt2;
                    unsigned todo = __ballot_sync(0xffffffffu, usable);
                    const unsigned alts = __ballot_sync(0xffffffffu, is_alt);
                    while (todo) {
                        const int leader = __ffs((int)todo) - 1;
                        const int name = __shfl_sync(0xffffffffu, amp, leader);
                        const unsigned same = __ballot_sync(0xffffffffu, usable && amp == name);
                        todo &= ~same;
                        if (lane == 0) {
                            int k = 0;
                            const int cnt = s_n[warp];
                            while (k < cnt && s_name[warp][k] != name) k++;
                            if (k == cnt) {
                                if (cnt == kAmpSlots) overflow = true;
                                else { s_name[warp][k] = name; s_cov[warp][k] = 0; s_sup[warp][k] = 0; s_n[warp] = cnt + 1; }
                            }
                            if (k < kAmpSlots) { s_cov[warp][k] += __popc(same); s_sup[warp][k] += __popc(same & alts); }
                        }
                    }
                }
            }
            prev_end = end;
        }
        __syncwarp();
        if (lane == 0) {
            if (overflow) { atomicExch(status, 1); continue; }
            const int cnt = s_n[warp];
            bool any_support = false;
            for (int k = 0; k < cnt; k++) any_support |= s_sup[warp][k] > 0;
            if (!any_support || cnt < 2) continue;                  // CalculateAmpliconBias :50-59: no verdict
            double max_freq = 0.0;
            for (int k = 0; k < cnt; k++) {                         // :64-82
                const double freq = (double)s_sup[warp][k] / (double)s_cov[warp][k];
                if (freq >= max_freq) max_freq = freq;
            }
            bool bias = false;
            for (int k = 0; k < cnt; k++) {                         // :84-129
                const double coverage = s_cov[warp][k], support = s_sup[warp][k];
                const double freq = support / coverage;
                const double expected = max_freq * coverage;
                double chance = 1.0;
                if (expected < 5.0) {}                               // Constants.MinNumObservations
                else if (expected <= support || freq > 0.1) {}       // Constants.FreePassObservationFreq
                else chance = fmax(0.0, pisces_poisson_cdf(support, expected));
                if (chance < (double)acceptance) bias = true;
            }
            if (bias) records[i].filters |= (uint16_t)(1u << FLT_AMPLICON_BIAS);
        }
        __syncwarp();
    }
}
}  // namespace

cudaError_t launch_pvert_count(const ReadsView& rv, const RegionView& rg, const int32_t* end_pos, int n_classes, int32_t* cls_rows, int32_t* complex, cudaStream_t st) {
    if (rv.n_reads == 0) return cudaSuccess;
    FillTargets ft{};
    return launch_walk<false>(rv, rg, end_pos, n_classes, cls_rows, ft, complex, -1, st);
}
cudaError_t launch_pvert_layout(int32_t* cls_rows, int32_t n_tiles, int n_classes, int64_t* tile_rows, cudaStream_t st) {
    pvert_layout_kernel<<<(n_tiles + 1 + 255) / 256, 256, 0, st>>>(cls_rows, n_tiles, n_classes, tile_rows);
    return cudaGetLastError();
}
cudaError_t launch_pvert_fill(const ReadsView& rv, const RegionView& rg, const int32_t* end_pos, int n_classes, const int64_t* tile_row0, const int32_t* cls_end, int32_t* cursor, uint8_t* data,
                              int2* row_meta, int32_t* row_amp, uint32_t* exc_entries, unsigned long long* exc_count, int64_t exc_capacity, const uint8_t* ref_slot,
                              int32_t* complex, int64_t n_complex, cudaStream_t st) {
    if (rv.n_reads == 0) return cudaSuccess;
    FillTargets ft{tile_row0, cls_end, cursor, data, row_meta, row_amp, exc_entries, exc_count, exc_capacity, reinterpret_cast<const uint32_t*>(ref_slot)};
    return launch_walk<true>(rv, rg, end_pos, n_classes, nullptr, ft, complex, n_complex, st);
}
cudaError_t launch_pvert_transpose(uint8_t* data, int64_t n_blocks, cudaStream_t st) {
    if (n_blocks <= 0) return cudaSuccess;
    pvert_transpose_kernel<<<(unsigned)((n_blocks + kTransposeWarps - 1) / kTransposeWarps), 32 * kTransposeWarps, 0, st>>>(data, n_blocks);
    return cudaGetLastError();
}
cudaError_t launch_pvert_ref_bases(const uint8_t* chr, int64_t chr_len, const int32_t* positions, int32_t first_position, int64_t n_loci, uint8_t* ref_base, uint8_t* ref_slot,
                                   cudaStream_t st) {
    if (n_loci <= 0) return cudaSuccess;
    const int64_t n_padded = ref_slot ? (n_loci + 31) / 32 * 32 : n_loci;
    pvert_ref_bases_kernel<<<(unsigned)((n_padded + 255) / 256), 256, 0, st>>>(chr, chr_len, positions, first_position, n_loci, ref_base, ref_slot, n_padded);
    return cudaGetLastError();
}
cudaError_t launch_pvert_gather(const PvertPileup& in, const int32_t* req_locus, int32_t n_req, int32_t* out_counts, int32_t* out_collapsed, int min_bq, cudaStream_t st) {
    if (n_req <= 0) return cudaSuccess;
    pvert_gather_kernel<<<(n_req + kPvGatherWarps - 1) / kPvGatherWarps, 32 * kPvGatherWarps, 0, st>>>(in, req_locus, n_req, out_counts, out_collapsed, min_bq);
    return cudaGetLastError();
}
cudaError_t launch_pvert_amplicon_bias(const PvertPileup& in, const int32_t* row_amp, pb2_call_record* records, const unsigned long long* n_dev, int64_t n_host, int64_t capacity,
                                       const uint8_t* valid, float acceptance, int min_bq, int* status, int num_sms, cudaStream_t st) {
    if (row_amp == nullptr || (n_dev == nullptr && n_host <= 0)) return cudaSuccess;
    const int64_t want = n_dev ? (int64_t)num_sms * 4 : (n_host + kAmpWarps - 1) / kAmpWarps;
    pvert_amplicon_bias_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms * 8)), 32 * kAmpWarps, 0, st>>>(in, row_amp, records, n_dev, n_host, capacity, valid,
                                                                                                                                      acceptance, min_bq, status);
    return cudaGetLastError();
}

}  // namespace pb2
