// Read expansion (K0): reads (struct of arrays) -> locus-major pileup entries in CSR form, on the device.
// One thread walks one read's CIGAR exactly as RegionStateManager.AddAlleleCounts does
// (src/lib/Pisces.Processing/RegionState/RegionStateManager.cs:118-220) and produces one entry per AddAlleleCount call; the SNV
// candidate flags follow CandidateVariantFinder (src/lib/Pisces.Domain/Logic/CandidateVariantFinder.cs:90-203,496-553) for CallMNVs=false.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "pb2_internal.hpp"

namespace pb2 {

__device__ __forceinline__ bool op_ref_span(int op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }   // M D N = X  (BamCommon.cs:560-573)
__device__ __forceinline__ bool op_read_span(int op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }  // M I S = X  (:575-588)

// RegionStateManager.GetAnchorType (:83-116), K = 5
__device__ __forceinline__ int anchor_type(int end_pos, int base_pos, int start_pos) {
    const int left = base_pos - start_pos, right = end_pos - base_pos;
    if (left >= right) return right >= kAnchorK ? kAnchorK : kNumAnchors - right - 1;
    return left >= kAnchorK ? kAnchorK : left;
}
// ReadExtentions.GetReadCollapsedType (Read.cs:17-64) from the per-read summary byte; returns type+1 or 0
__device__ __forceinline__ int collapsed_code(int cbyte, int dir) {
    if (!(cbyte & 1)) return 0;
    if (cbyte & 2) return (dir == DIR_S ? 0 : 1) + 1;                // DuplexStitched / DuplexNonStitched
    const int pd = (cbyte >> 2) & 3;
    if (pd == 1) return (dir == DIR_S ? 4 : 5) + 1;                  // SimplexForward(Non)Stitched
    if (pd == 2) return (dir == DIR_S ? 6 : 7) + 1;                  // SimplexReverse(Non)Stitched
    return 0;
}

// Walks one read; calls emit(position, code, qual, anchor_byte) once per pileup entry.
template <class Emit>
__device__ void walk_read(const ReadsView& rv, const RegionView& rg, int r, Emit emit) {
    const int64_t c0 = rv.cigar_off[r], c1 = rv.cigar_off[r + 1];
    const int64_t s0 = rv.seq_off[r];
    const int read_len = (int)(rv.seq_off[r + 1] - s0);
    const int n_ops = (int)(c1 - c0);
    if (n_ops == 0) return;
    const int start_pos = rv.pos0[r] + 1;                                  // Read.Position
    int ref_span = 0;
    for (int i = 0; i < n_ops; i++) { const uint32_t c = rv.cigar[c0 + i]; if (op_ref_span(c & 15)) ref_span += (int)(c >> 4); }
    const int end_pos = rv.pos0[r] + ref_span;                             // Read.EndPosition (Read.cs:88-91)
    const bool reverse = (rv.flag[r] & 0x10) != 0;
    const int cbyte = (rg.expect_collapsed && rv.collapsed) ? rv.collapsed[r] : 0;
    const uint8_t* bases = rv.bases + s0;
    const uint8_t* quals = rv.quals + s0;
    const uint8_t* dirs = rv.base_dirs ? rv.base_dirs + s0 : nullptr;
    auto dir_at = [&](int i) -> int { return dirs ? dirs[i] : (reverse ? DIR_R : DIR_F); };
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
    // open-end annotation (:496-553): first / last non-soft-clip operation
    int first_op = (int)(rv.cigar[c0] & 15);
    if (first_op == 4 && n_ops >= 2) first_op = (int)(rv.cigar[c0 + 1] & 15);
    int last_nonclip = last_op;
    if (last_nonclip == 4 && n_ops >= 2) last_nonclip = prev_op;
    int max_mapped = -1;
    {
        int rp = start_pos;
        for (int i = 0; i < n_ops; i++) {
            const uint32_t c = rv.cigar[c0 + i];
            const int op = c & 15, len = (int)(c >> 4);
            if (op_ref_span(op)) { if (op_read_span(op) && len > 0) max_mapped = rp + len - 1; rp += len; }
        }
    }

    int read_idx = 0, ref_pos = start_pos, last_position = start_pos - 1;
    for (int oi = 0; oi < n_ops; oi++) {
        const uint32_t c = rv.cigar[c0 + oi];
        const int op = c & 15, len = (int)(c >> 4);
        const bool rs = op_read_span(op), fs = op_ref_span(op);
        if (rs) {
            for (int k = 0; k < len; k++, read_idx++) {
                const int dir = dir_at(read_idx);
                if (ends_in_del_before_clip && read_idx == len_before_del) {      // (:148-159)
                    const int dq = del_q(read_idx);
                    for (int j = 1; j < del_len + 1; j++)
                        emit(j + last_position, AT_DEL | (dir << 3), dq, (kNumAnchors - 1) | (collapsed_code(cbyte, dir) << 4));
                }
                if (!fs) continue;                                               // I / S: not mapped to the reference
                const int position = ref_pos++;
                const int an = anchor_type(end_pos, position, start_pos);
                const int cc = collapsed_code(cbyte, dir) << 4;
                if (position > last_position + 1) {                              // deletion (or N skip) before this base (:170-177)
                    const int dq = del_q(read_idx);
                    for (int j = last_position + 1; j < position; j++) emit(j, AT_DEL | (dir << 3), dq, an | cc);
                }
                const uint8_t b = bases[read_idx];
                const int allele = b == 'A' ? AT_A : b == 'C' ? AT_C : b == 'G' ? AT_G : b == 'T' ? AT_T : AT_N;
                int code = allele | (dir << 3);
                // SNV candidate flags for CallMNVs=false (CandidateVariantFinder.cs:90-168): only 'M' operations inside the chromosome raise candidates
                const bool in_chr = rg.chr == nullptr || position <= rg.chr_len;
                if (op != 0 || !in_chr) code |= PB2_ENTRY_NO_CANDIDATE;
                else {
                    // open on the right: the next base of this operation exists and is unusable (low quality, N, or reference N) -> FlushVariant(..., openRight=true)
                    if (k + 1 < len) {
                        const int nq = quals[read_idx + 1];
                        const uint8_t nb = bases[read_idx + 1];
                        const bool nb_n = !(nb == 'A' || nb == 'C' || nb == 'G' || nb == 'T');
                        bool nref_n = false, n_in = true;
                        if (rg.chr) {
                            n_in = position + 1 <= rg.chr_len;
                            if (n_in) { const uint8_t rb = rg.chr[position]; nref_n = !(rb == 'A' || rb == 'C' || rb == 'G' || rb == 'T'); }
                        }
                        if (n_in && (nq < rg.min_bq || nb_n || nref_n)) code |= PB2_ENTRY_OPEN_RIGHT;
                    }
                    if (first_op == 0 && position == start_pos) code |= PB2_ENTRY_OPEN_LEFT;
                    if (last_nonclip == 0 && position == max_mapped) code |= PB2_ENTRY_OPEN_RIGHT;
                }
                emit(position, code, quals[read_idx], an | cc);
                last_position = position;
            }
        } else if (fs) {
            ref_pos += len;
        }
    }
    if (ends_in_del) {                                                           // (:195-210)
        const int dq = del_q(read_len - 1);
        const int dir = read_len > 0 ? dir_at(read_len - 1) : DIR_F;
        if (read_len > 0)
            for (int j = 1; j < del_len + 1; j++) emit(j + last_position, AT_DEL | (dir << 3), dq, (kNumAnchors - 1) | (collapsed_code(cbyte, dir) << 4));
    }
}

__device__ __forceinline__ int64_t locus_index(const RegionView& rg, int position) {
    if (position < rg.lo || position > rg.hi) return -1;
    return rg.index_of_pos ? (int64_t)rg.index_of_pos[position - rg.lo] : (int64_t)(position - rg.lo);
}

__global__ void reads_count_kernel(ReadsView rv, RegionView rg, unsigned int* __restrict__ depth) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rv.n_reads) return;
    walk_read(rv, rg, r, [&](int position, int, int, int) {
        const int64_t li = locus_index(rg, position);
        if (li >= 0) atomicAdd(depth + li, 1u);
    });
}
__global__ void reads_emit_kernel(ReadsView rv, RegionView rg, const int64_t* __restrict__ offsets, unsigned int* __restrict__ cursor, uint8_t* __restrict__ code,
                                  uint8_t* __restrict__ qual, uint8_t* __restrict__ anch) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rv.n_reads) return;
    walk_read(rv, rg, r, [&](int position, int c, int q, int a) {
        const int64_t li = locus_index(rg, position);
        if (li < 0) return;
        const int64_t o = offsets[li] + (int64_t)atomicAdd(cursor + li, 1u);
        code[o] = (uint8_t)c;
        qual[o] = (uint8_t)max(q, 0);
        anch[o] = (uint8_t)a;
    });
}
__global__ void depth_to_i64_kernel(const unsigned int* __restrict__ depth, int64_t* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = depth[i]; else if (i == n) out[i] = 0;
}

// ------------------------------------------------------------------------------------------------ candidate finder
// CandidateVariantFinder.FindCandidates (CandidateVariantFinder.cs:31-83) for one read per thread: insertions (:234-260), deletions (:262-292) and —
// with CallMNVs — the SNV/MNV state machine over 'M' operations (:90-203); Create (:334-387: support direction, well-anchored support, collapsed-read
// type) and the open-end annotation (:496-553). Each candidate found becomes one RawCand; the host sums them per key (RegionState.AddCandidate).
// Without CallMNVs the SNV candidates are the pileup counts themselves (flag bits of the entries), so 'M' operations emit nothing here.
__device__ __forceinline__ void emit_raw(RawCand* out, unsigned long long* count, int64_t capacity, const RawCand& rc) {
    const unsigned long long slot = atomicAdd(count, 1ull);
    if ((int64_t)slot < capacity) out[slot] = rc;
}

__global__ void reads_candidates_kernel(ReadsView rv, int32_t first_read, const uint8_t* __restrict__ chr, int64_t chr_len, int min_bq, int call_mnvs, int max_mnv,
                                        int max_gap, int expect_collapsed, RawCand* __restrict__ out, unsigned long long* __restrict__ count, int64_t capacity,
                                        int32_t pos_lo, int32_t pos_hi, int snv_only, const int32_t* __restrict__ end_pos_of) {
    const int r = first_read + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rv.n_reads) return;
    // only candidates at positions in (pos_lo, pos_hi] are wanted: reads that cannot hold one are skipped (an insertion before the first base sits at Position - 1)
    if (rv.pos0[r] > pos_hi || (end_pos_of != nullptr && end_pos_of[r] < pos_lo)) return;
    const int64_t c0 = rv.cigar_off[r], c1 = rv.cigar_off[r + 1];
    const int64_t s0 = rv.seq_off[r];
    const int read_len = (int)(rv.seq_off[r + 1] - s0);
    const int n_ops = (int)(c1 - c0);
    if (n_ops == 0 || read_len == 0 || chr == nullptr) return;
    const int start_pos = rv.pos0[r] + 1;
    int ref_span = 0, max_mapped = -1;
    {
        int rp = start_pos;
        for (int i = 0; i < n_ops; i++) {
            const uint32_t c = rv.cigar[c0 + i];
            const int op = c & 15, len = (int)(c >> 4);
            if (op_ref_span(op)) { ref_span += len; if (op_read_span(op) && len > 0) max_mapped = rp + len - 1; rp += len; }
        }
    }
    const int end_pos = rv.pos0[r] + ref_span;                         // Read.EndPosition
    if (max_mapped == -1) max_mapped = start_pos - 1;                  // (:506-507)
    const bool reverse = (rv.flag[r] & 0x10) != 0;
    const int cbyte = (expect_collapsed && rv.collapsed) ? rv.collapsed[r] : 0;
    const uint8_t* bases = rv.bases + s0;
    const uint8_t* quals = rv.quals + s0;
    const uint8_t* dirs = rv.base_dirs ? rv.base_dirs + s0 : nullptr;
    auto dir_at = [&](int i) -> int { return dirs ? dirs[i] : (reverse ? DIR_R : DIR_F); };
    int first_op = (int)(rv.cigar[c0] & 15), last_op = (int)(rv.cigar[c1 - 1] & 15);
    if (first_op == 4 && n_ops >= 2) first_op = (int)(rv.cigar[c0 + 1] & 15);
    if (last_op == 4 && n_ops >= 2) last_op = (int)(rv.cigar[c1 - 2] & 15);

    // GetSupportDirection (:396-445); deletions in stitched reads use the per-base directions on both sides of the gap
    auto support_dir = [&](int type, int start_idx, int length) -> int {
        if (type == CAT_SNV) return dir_at(start_idx);
        const int left = start_idx - 1;
        const int right = type == CAT_DEL ? start_idx : start_idx + length;
        const int last = read_len - 1;
        if (right == 0) return dir_at(0);
        if (left == last) return dir_at(last);
        if (left == right - 1) { const int s = dir_at(left), e = dir_at(right); return s == DIR_S ? e : s; }
        int d = DIR_F;
        for (int i = left + 1; i < right && i < read_len; i++) { d = dir_at(i); if (d == DIR_S) return DIR_S; }
        return d;
    };
    auto create = [&](int type, int position, int ref_len, int alt_len, int start_idx, int order, bool open_l, bool open_r) {
        const int length = type == CAT_INS ? alt_len - 1 : (type == CAT_DEL ? ref_len - 1 : alt_len);
        const int d = support_dir(type, start_idx, length);
        const int anchor = min(position - start_pos, end_pos - position);
        // Annotate (:496-553)
        if (first_op == 0 && position == start_pos && (type == CAT_MNV || type == CAT_SNV)) open_l = true;
        if (first_op == 1 && position == start_pos - 1 && type == CAT_INS) open_l = true;
        if (first_op == 2 && position == start_pos - 1 && type == CAT_DEL) open_l = true;
        if (last_op == 0 && position + alt_len - 1 == max_mapped && (type == CAT_MNV || type == CAT_SNV)) open_r = true;
        if (last_op == 1 && position == max_mapped && type == CAT_INS) open_r = true;
        if (last_op == 2 && position == max_mapped && type == CAT_DEL) open_r = true;
        RawCand rc;
        rc.read = r; rc.order = order; rc.position = position; rc.start_in_read = start_idx;
        rc.ref_len = (uint16_t)ref_len; rc.alt_len = (uint16_t)alt_len;
        rc.type = (uint8_t)type; rc.dir = (uint8_t)d;
        rc.flags = (uint8_t)((open_l ? 1 : 0) | (open_r ? 2 : 0) | (anchor > min(kAnchorK - 1, alt_len - 1) ? 4 : 0));
        rc.collapsed = (uint8_t)collapsed_code(cbyte, d);
        if (position <= pos_lo || position > pos_hi || (snv_only && type != CAT_SNV)) return;
        const int n_from_read = type == CAT_INS ? alt_len - 1 : (type == CAT_DEL ? 0 : alt_len);
#pragma unroll
        for (int k = 0; k < 8; k++) rc.read_bases[k] = (k < n_from_read && start_idx + k < read_len) ? bases[start_idx + k] : (uint8_t)0;
        emit_raw(out, count, capacity, rc);
    };
    auto is_n = [](uint8_t b) { return !(b == 'A' || b == 'C' || b == 'G' || b == 'T'); };

    int read_idx = 0, ref_idx = start_pos - 1;   // startIndexInRead / startIndexInReference (0-based)
    for (int oi = 0; oi < n_ops; oi++) {
        const uint32_t c = rv.cigar[c0 + oi];
        const int op = c & 15, len = (int)(c >> 4);
        if (op == 1) {                                                  // insertion (:234-260)
            if (!(ref_idx - 1 >= chr_len || ref_idx == 0) && (int)quals[read_idx] >= min_bq) create(CAT_INS, ref_idx, 1, 1 + len, read_idx, oi << 16, false, false);
        } else if (op == 2) {                                           // deletion (:262-292)
            if (!((int64_t)ref_idx + len >= chr_len) && ref_idx >= 1) {
                const int after = read_idx < read_len ? quals[read_idx] : quals[read_idx - 1];
                const int before = read_idx > 0 ? quals[read_idx - 1] : after;
                if (before >= min_bq && after >= min_bq) create(CAT_DEL, ref_idx, len + 1, 1, read_idx, oi << 16, false, false);
            }
        } else if (op == 0 && call_mnvs) {                              // ExtractSnvsFromOperation (:90-168)
            int vlen = 0, gap = 0;
            bool open_left = false;
            auto flush = [&](int i_end, bool open_right) {             // FlushVariant (:183-203); the variant ends just before operation offset i_end
                int v = vlen;
                if (gap >= 1) { v -= gap; open_right = false; }
                if (v >= 1) {
                    const int st = i_end - vlen;                        // offset of the variant's first base in the operation
                    create(v > 1 ? CAT_MNV : CAT_SNV, ref_idx + st + 1, v, v, read_idx + st, (oi << 16) | min(st, 0xffff), open_left, open_right);
                }
            };
            auto should_build = [&](bool ref_next) {                   // ShouldBuildUpMNV (:170-181)
                if (ref_next && vlen == 0) return false;
                if (vlen + 1 > max_mnv) return false;
                if (gap + (ref_next ? 1 : 0) > max_gap) return false;
                return true;
            };
            int i = 0;
            for (; i < len; i++) {
                if ((int64_t)ref_idx + i >= chr_len) break;
                const bool good = (int)quals[read_idx + i] >= min_bq;
                const uint8_t rb = bases[read_idx + i], fb = chr[ref_idx + i];
                const bool at_end = i == len - 1;
                const bool starting_at_end = at_end && vlen == 0;
                if (is_n(rb) || is_n(fb) || !good) {
                    flush(i, true);
                    vlen = 0; gap = 0; open_left = true;
                } else if (fb == rb) {
                    if (should_build(true) && !starting_at_end) { vlen++; gap++; }
                    else { flush(i, false); vlen = 0; gap = 0; open_left = false; }
                } else {
                    if (should_build(false) && !starting_at_end) { vlen++; gap = 0; }
                    else { flush(i, false); vlen = 1; gap = 0; open_left = false; }
                }
            }
            // the final flush of the operation uses operationLength even when the loop stopped at the chromosome end (:166-167)
            flush(len, false);
        }
        if (op_read_span(op)) read_idx += len;
        if (op_ref_span(op)) ref_idx += len;
    }
}

// ------------------------------------------------------------------------------------------------ the device read store (pb2_push_reads)
__global__ void reads_ingest_kernel(ReadsView rv, int32_t first_read, int64_t cigar_base, int64_t seq_base, int64_t* __restrict__ cigar_off, int64_t* __restrict__ seq_off,
                                    int32_t* __restrict__ end_pos, int32_t prev_key, int2* __restrict__ triggers, int32_t trigger_capacity, IngestStatus* __restrict__ status,
                                    int expect_collapsed) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = rv.n_reads - first_read;
    int lo = INT32_MAX, hi = 0;
    if (j < nb) {
        const int r = first_read + j;
        // offsets as pushed (relative to the batch); the rebased values are written by the thread of the read they start (and thread nb - 1 the last one)
        const int64_t c0 = cigar_off[r] + (j == 0 ? 0 : cigar_base), c1 = cigar_off[r + 1] + cigar_base;
        const int64_t s0 = seq_off[r] + (j == 0 ? 0 : seq_base), s1 = seq_off[r + 1] + seq_base;
        int err = 0;
        if (c1 < c0 || s1 < s0) err = 4;
        else if (rv.pos0[r] < 0) err = 3;
        int64_t rs = 0, fs = 0;
        if (!err) {
            for (int64_t k = c0; k < c1; k++) {
                const uint32_t c = rv.cigar[k];
                const int op = c & 15;
                if (op > 8) { err = 1; break; }
                if ((op == 1 || op == 2) && (c >> 4) > 65534u) { err = 6; break; }   // candidate allele lengths travel as 16-bit fields
                if (op_read_span(op)) rs += c >> 4;
                if (op_ref_span(op)) fs += c >> 4;
            }
            if (!err && c1 > c0 && rs != s1 - s0) err = 2;
            // CollapsedRegionStateManager.AddCollapsedReadCount (CollapedRegionStateManager.cs:40-43) throws on the first counted base of a read without XV / XW
            if (!err && expect_collapsed && c1 > c0 && fs > 0 && !(rv.collapsed != nullptr && (rv.collapsed[r] & 1))) err = 5;
        }
        if (err) { if (atomicCAS(&status->error, 0, err) == 0) status->error_read = j; }
        const int e = rv.pos0[r] + (int)fs;
        end_pos[r] = e;
        if (!err) { lo = rv.pos0[r] + 1; hi = e; }
        const int up_to = rv.pos0[r];
        const int key = up_to <= 0 ? 0 : (up_to + 999) / 1000;
        int pk = prev_key;
        if (j > 0) { const int pp = rv.pos0[r - 1]; pk = pp <= 0 ? 0 : (pp + 999) / 1000; }
        if (key != pk) {
            const int slot = atomicAdd(&status->n_triggers, 1);
            if (slot < trigger_capacity) triggers[slot] = make_int2(r, up_to);
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { if (lo != INT32_MAX) atomicMin(&status->min_start, lo); if (hi > 0) atomicMax(&status->max_end, hi); }
}
__global__ void reads_rebase_kernel(int32_t first_read, int32_t nb, int64_t cigar_base, int64_t seq_base, int64_t* __restrict__ cigar_off, int64_t* __restrict__ seq_off) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < 1 || j > nb) return;   // entry `first_read` already holds the absolute start of the batch
    cigar_off[first_read + j] += cigar_base;
    seq_off[first_read + j] += seq_base;
}
cudaError_t launch_reads_ingest(const ReadsView& rv, int32_t first_read, int64_t cigar_base, int64_t seq_base, int64_t* cigar_off, int64_t* seq_off, int32_t* end_pos,
                                int32_t prev_key, int2* triggers, int32_t trigger_capacity, IngestStatus* status, int expect_collapsed, cudaStream_t st) {
    const int nb = rv.n_reads - first_read;
    if (nb <= 0) return cudaSuccess;
    reads_ingest_kernel<<<(nb + 127) / 128, 128, 0, st>>>(rv, first_read, cigar_base, seq_base, cigar_off, seq_off, end_pos, prev_key, triggers, trigger_capacity, status,
                                                         expect_collapsed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    reads_rebase_kernel<<<(nb + 1 + 255) / 256, 256, 0, st>>>(first_read, nb, cigar_base, seq_base, cigar_off, seq_off);
    return cudaGetLastError();
}

// slot = allele2 << 6 | quality clamped to [1, 63]; a base that is not A/C/G/T is (0, quality 1): counted, never an allele (pb2_pvert.cuh)
__device__ __forceinline__ uint32_t slots_of_word(uint32_t b4, uint32_t q4) {
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t b = (b4 >> (8 * k)) & 0xffu, q = (q4 >> (8 * k)) & 0xffu;
        const uint32_t d = b - 65u;                                                     // 'A'
        const bool valid = d < 20u && ((0x80045u >> d) & 1u);                           // A C G T
        const uint32_t a2 = (0x78u >> (2u * ((b >> 1) & 3u))) & 3u;                     // A 0, C 2, G 1, T 3 (AlleleType order A G C T)
        out |= (valid ? ((a2 << 6) | min(max(q, 1u), 63u)) : 1u) << (8 * k);
    }
    return out;
}
__global__ void reads_slots_kernel(const uint8_t* __restrict__ bases, const uint8_t* __restrict__ quals, int64_t n, uint8_t* __restrict__ slots) {
    // 16 bytes per thread where the three pointers allow it (they share their offset into 256-byte aligned planes, the slot plane 16 bytes further)
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i >= n) return;
    const bool aligned = ((reinterpret_cast<uintptr_t>(bases + i) | reinterpret_cast<uintptr_t>(quals + i) | reinterpret_cast<uintptr_t>(slots + i)) & 15u) == 0;
    if (aligned && i + 16 <= n) {
        const uint4 b = *reinterpret_cast<const uint4*>(bases + i), q = *reinterpret_cast<const uint4*>(quals + i);
        *reinterpret_cast<uint4*>(slots + i) = make_uint4(slots_of_word(b.x, q.x), slots_of_word(b.y, q.y), slots_of_word(b.z, q.z), slots_of_word(b.w, q.w));
    } else {
        for (int64_t k = i; k < n && k < i + 16; k++) slots[k] = (uint8_t)slots_of_word(bases[k], quals[k]);
    }
}
cudaError_t launch_reads_slots(const uint8_t* bases, const uint8_t* quals, int64_t n, uint8_t* slots, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    reads_slots_kernel<<<(unsigned)(((n + 15) / 16 + 255) / 256), 256, 0, st>>>(bases, quals, n, slots);
    return cudaGetLastError();
}

__global__ void reads_keep_flags_kernel(const int32_t* __restrict__ end_pos, int64_t n, int32_t cleared_to, const int64_t* __restrict__ cigar_off,
                                        const int64_t* __restrict__ seq_off, int64_t* __restrict__ keep_reads, int64_t* __restrict__ keep_cigar, int64_t* __restrict__ keep_seq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const bool keep = i < n && end_pos[i] > cleared_to;
    keep_reads[i] = keep ? 1 : 0;
    keep_cigar[i] = keep ? cigar_off[i + 1] - cigar_off[i] : 0;
    keep_seq[i] = keep ? seq_off[i + 1] - seq_off[i] : 0;
}
cudaError_t launch_reads_keep_flags(const int32_t* end_pos, int64_t n, int32_t cleared_to, const int64_t* cigar_off, const int64_t* seq_off, int64_t* keep_reads,
                                    int64_t* keep_cigar, int64_t* keep_seq, cudaStream_t st) {
    reads_keep_flags_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(end_pos, n, cleared_to, cigar_off, seq_off, keep_reads, keep_cigar, keep_seq);
    return cudaGetLastError();
}
// one warp per read: the kept reads move to their new places (exclusive scans of the keep flags / lengths)
__global__ void reads_compact_kernel(ReadsCompactArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= a.n) return;
    const int64_t ni = a.new_index[i];
    if (a.new_index[i + 1] == ni) return;   // dropped
    const int64_t c0 = a.cigar_off[i], nc = a.cigar_off[i + 1] - c0, s0 = a.seq_off[i], ns = a.seq_off[i + 1] - s0;
    const int64_t oc = a.new_cigar[i], os = a.new_seq[i];
    if (lane == 0) {
        a.o_pos0[ni] = a.pos0[i]; a.o_end_pos[ni] = a.end_pos[i]; a.o_flag[ni] = a.flag[i];
        a.o_cigar_off[ni] = oc; a.o_seq_off[ni] = os;
        if (a.collapsed) a.o_collapsed[ni] = a.collapsed[i];
        if (a.amplicon) a.o_amplicon[ni] = a.amplicon[i];
        if (a.new_index[a.n] == ni + 1) { a.o_cigar_off[ni + 1] = oc + nc; a.o_seq_off[ni + 1] = os + ns; }   // the last kept read closes the offset arrays
    }
    for (int64_t k = lane; k < nc; k += 32) a.o_cigar[oc + k] = a.cigar[c0 + k];
    for (int64_t k = lane; k < ns; k += 32) {
        a.o_bases[os + k] = a.bases[s0 + k];
        a.o_quals[os + k] = a.quals[s0 + k];
        if (a.base_dirs) a.o_base_dirs[os + k] = a.base_dirs[s0 + k];
        if (a.slots) a.o_slots[os + k] = a.slots[s0 + k];
    }
}
cudaError_t launch_reads_compact(const ReadsCompactArgs& a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    reads_compact_kernel<<<(unsigned)((a.n * 32 + 255) / 256), 256, 0, st>>>(a);
    return cudaGetLastError();
}
__global__ void reads_block_bitmap_kernel(const int32_t* __restrict__ pos0, const int32_t* __restrict__ end_pos, int64_t n, int32_t cleared_through, int32_t key0,
                                          int32_t n_keys, uint32_t* __restrict__ bitmap) {
    // reads arrive in position order: the 32 reads of a warp touch a handful of neighbouring blocks. The warp's key range is reduced first and its
    // lanes then set the bits of that range (usually one bit, set by one lane)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k_lo = INT32_MAX, k_hi = INT32_MIN;
    if (i < n) {
        const int a = max(pos0[i] + 1, cleared_through + 1), e = end_pos[i];
        if (a <= e) { k_lo = (a + 999) / 1000; k_hi = (e + 999) / 1000; }
    }
    k_lo = __reduce_min_sync(0xffffffffu, k_lo);
    k_hi = __reduce_max_sync(0xffffffffu, k_hi);
    if (k_lo > k_hi) return;
    // (a warp whose reads are far apart would mark blocks between them that no read touches: only when the range is one every read of the warp
    // overlaps - the common case - is it used; otherwise every lane marks its own blocks)
    const int lane = threadIdx.x & 31;
    bool mine = false;
    int m_lo = 0, m_hi = -1;
    if (i < n) {
        const int a = max(pos0[i] + 1, cleared_through + 1), e = end_pos[i];
        if (a <= e) { m_lo = (a + 999) / 1000; m_hi = (e + 999) / 1000; mine = true; }
    }
    const bool same = !mine || (m_lo == k_lo && m_hi == k_hi);
    if (__all_sync(0xffffffffu, same)) {
        for (int k = k_lo + lane; k <= k_hi; k += 32) { const int b = k - key0; if (b >= 0 && b < n_keys) atomicOr(bitmap + (b >> 5), 1u << (b & 31)); }
    } else {
        for (int k = m_lo; k <= m_hi; k++) { const int b = k - key0; if (b >= 0 && b < n_keys) atomicOr(bitmap + (b >> 5), 1u << (b & 31)); }
    }
}
cudaError_t launch_reads_block_bitmap(const int32_t* pos0, const int32_t* end_pos, int64_t n, int32_t cleared_through, int32_t key0, int32_t n_keys, uint32_t* bitmap,
                                      cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    reads_block_bitmap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pos0, end_pos, n, cleared_through, key0, n_keys, bitmap);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ RegionState.AddCandidate on the device (reduce by key)
// The candidates a read set raises are many (every read that carries an indel raises it again) and few distinct. They are grouped here: sorted by
// (position, hash of the allele), equal neighbours merged with their counts summed - what RegionState.AddCandidate (RegionState.cs:94-174) does one
// candidate at a time - and the host receives one row per distinct candidate with the (read, order) of its first occurrence, which is all it needs to
// insert them in the reference's order. Two different candidates that share position and hash would be merged wrongly: the head test compares all
// fields and raises `collision`, and the host then takes the one-by-one path instead.
__device__ __forceinline__ bool raw_same(const RawCand& a, const RawCand& b) {
    if (a.position != b.position || a.type != b.type || (a.flags & 3) != (b.flags & 3) || a.ref_len != b.ref_len || a.alt_len != b.alt_len) return false;
#pragma unroll
    for (int k = 0; k < 8; k++) if (a.read_bases[k] != b.read_bases[k]) return false;
    return true;
}
__global__ void cand_keys_kernel(const RawCand* __restrict__ raw, int64_t n, unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx, int32_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RawCand c = raw[i];
    uint32_t hsh = 2166136261u;
    auto mix = [&](uint32_t v) { hsh = (hsh ^ v) * 16777619u; };
    mix(c.type); mix(c.flags & 3u); mix(c.ref_len); mix(c.alt_len);
#pragma unroll
    for (int k = 0; k < 8; k++) mix(c.read_bases[k]);
    keys[i] = ((unsigned long long)(uint32_t)c.position << 32) | hsh;
    idx[i] = (uint32_t)i;
    const int n_from_read = c.type == CAT_INS ? (int)c.alt_len - 1 : (c.type == CAT_DEL ? 0 : (int)c.alt_len);
    if (n_from_read > 8) atomicOr(flags, 1);   // alleles longer than RawCand::read_bases cannot be compared here
}
__global__ void cand_heads_kernel(const RawCand* __restrict__ raw, const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ idx, int64_t n,
                                  int32_t* __restrict__ head, int32_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int hd = 1;
    if (i > 0 && keys[i] == keys[i - 1]) {
        hd = 0;
        if (!raw_same(raw[idx[i]], raw[idx[i - 1]])) atomicOr(flags, 2);
    }
    head[i] = hd;
}
__global__ void cand_reduce_kernel(const RawCand* __restrict__ raw, const uint32_t* __restrict__ idx, const int32_t* __restrict__ group_of /* inclusive scan of head */,
                                   int64_t n, CandGroup* __restrict__ groups) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RawCand c = raw[idx[i]];
    CandGroup& g = groups[group_of[i] - 1];
    atomicAdd(&g.support[c.dir], 1);
    if (c.flags & 4) atomicAdd(&g.well_anchored[c.dir], 1);
    if (c.collapsed) {   // CandidateVariantFinder.Create (:352-384)
        const int t = c.collapsed - 1;
        atomicAdd(&g.collapsed_mut[t], 1);
        if (t == 4 || t == 6) atomicAdd(&g.collapsed_mut[2], 1);
        else if (t == 5 || t == 7) atomicAdd(&g.collapsed_mut[3], 1);
    }
    const unsigned long long seen = ((unsigned long long)(uint32_t)c.read << 32) | (uint32_t)c.order;
    const unsigned long long old = atomicMin(&g.first_seen, seen);
    if (seen < old) g.first_index = idx[i];   // (racy among equal candidates only in which of them is kept: they are equal in every field that is read)
}
__global__ void cand_groups_init_kernel(CandGroup* __restrict__ groups, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    CandGroup g;
    memset(&g, 0, sizeof(g));
    g.first_seen = ~0ull;
    groups[i] = g;
}
cudaError_t cand_reduce_temp_bytes(int64_t n, size_t* bytes) {
    size_t a = 0, b = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::InclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
    *bytes = std::max(a, b);
    return e;
}
cudaError_t launch_cand_group(const RawCand* raw, int64_t n, unsigned long long* keys_in, unsigned long long* keys_out, uint32_t* idx_in, uint32_t* idx_out, int32_t* head,
                              int32_t* group_of, void* temp, size_t temp_bytes, int32_t* flags, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    cand_keys_kernel<<<grid, 256, 0, st>>>(raw, n, keys_in, idx_in, flags);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, idx_out, (int)n, 0, 64, st);
    if (e != cudaSuccess) return e;
    cand_heads_kernel<<<grid, 256, 0, st>>>(raw, keys_out, idx_out, n, head, flags);
    e = cub::DeviceScan::InclusiveSum(temp, temp_bytes, head, group_of, (int)n, st);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}
cudaError_t launch_cand_reduce(const RawCand* raw, const uint32_t* idx, const int32_t* group_of, int64_t n, CandGroup* groups, int64_t n_groups, cudaStream_t st) {
    cand_groups_init_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(groups, n_groups);
    cand_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(raw, idx, group_of, n, groups);
    return cudaGetLastError();
}

cudaError_t launch_reads_candidates(const ReadsView& rv, int32_t first_read, const uint8_t* chr, int64_t chr_len, int min_bq, int call_mnvs, int max_mnv, int max_gap,
                                    int expect_collapsed, RawCand* out, unsigned long long* count, int64_t capacity, int32_t pos_lo, int32_t pos_hi, int snv_only,
                                    const int32_t* end_pos, cudaStream_t st) {
    const int n = rv.n_reads - first_read;
    if (n <= 0) return cudaSuccess;
    reads_candidates_kernel<<<(n + 127) / 128, 128, 0, st>>>(rv, first_read, chr, chr_len, min_bq, call_mnvs, max_mnv, max_gap, expect_collapsed, out, count, capacity, pos_lo,
                                                            pos_hi, snv_only, end_pos);
    return cudaGetLastError();
}

cudaError_t launch_reads_count(const ReadsView& rv, const RegionView& rg, unsigned int* depth, cudaStream_t st) {
    if (rv.n_reads == 0) return cudaSuccess;
    reads_count_kernel<<<(rv.n_reads + 127) / 128, 128, 0, st>>>(rv, rg, depth);
    return cudaGetLastError();
}
cudaError_t launch_reads_emit(const ReadsView& rv, const RegionView& rg, const int64_t* offsets, unsigned int* cursor, uint8_t* code, uint8_t* qual, uint8_t* anch,
                              cudaStream_t st) {
    if (rv.n_reads == 0) return cudaSuccess;
    reads_emit_kernel<<<(rv.n_reads + 127) / 128, 128, 0, st>>>(rv, rg, offsets, cursor, code, qual, anch);
    return cudaGetLastError();
}
cudaError_t launch_depth_to_i64(const unsigned int* depth, int64_t* out, int64_t n, cudaStream_t st) {
    depth_to_i64_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(depth, out, n);
    return cudaGetLastError();
}

}  // namespace pb2
