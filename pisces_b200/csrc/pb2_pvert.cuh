// "PVERT": the read-major, bit-sliced pileup that pb2_push_reads builds on the device and the hot kernel counts with POPC.
//
// A tile is 32 consecutive loci (lane = locus). Every aligned stretch of a read inside a tile is a ROW of 32 one-byte slots
//     slot = 0                      no entry of this read at this locus
//     slot = allele2 << 6 | q6      a base: allele2 = AlleleType A 0, G 1, C 2, T 3; q6 = base quality clamped to [1, 63]; an N base is (0, q = 1)
//     slot = q6                     (deletion rows) one Deletion entry whose quality min(flanks) passed the bar (CandidateVariantFinder.cs:294-320)
// so a row is exactly what RegionStateManager.AddAlleleCounts (RegionStateManager.cs:118-220) would add for that stretch, before the quality rule.
// Direction, collapsed-read category and entry kind are properties of the read (stretch), not of the base: rows are grouped into CLASSES
//     class = kind + 2 * (direction + 3 * collapsed group)    kind 0 base / 1 deletion; collapsed group 0 none, 1 duplex, 2 simplex FR, 3 simplex RF
// and the rows of one class of one tile are stored together, padded to a multiple of 32 rows. 32 rows x 32 loci are then bit-transposed into a
// 1 KB BLOCK: for lane l eight 32-bit words B0 B1 Q5 Q4 | Q3 Q2 Q1 Q0 whose bit r is that bit of row r's slot at locus l (two halves of 512 B, each
// [lane][4 words], so that a warp reads a block with two fully coalesced 16-byte loads per lane). One byte per slot, no per-entry direction /
// anchor / collapsed bytes: the anchor bin of an entry follows from its row's read (row_meta), only the explicit-candidate gather needs it.
//
// The hot kernel needs min_base_call_quality in [2, 63] for this form (q = 0 / 1 are indistinguishable from "low quality", 0 marks an empty slot);
// quality sums (Window noise model) need every entry's own quality in reference order: those configurations keep the PTILE32 form.
#pragma once
#include "pb2_kernels.cuh"

namespace pb2 {

constexpr int kPvClassesPlain = 2 * 3;
constexpr int kPvClassesCollapsed = 2 * 3 * 4;
constexpr int kPvMaxClasses = kPvClassesCollapsed;
__host__ __device__ inline int pv_class(int kind, int dir, int cg) { return kind + 2 * (dir + 3 * cg); }
__host__ __device__ inline int pv_class_kind(int c) { return c & 1; }
__host__ __device__ inline int pv_class_dir(int c) { return (c >> 1) % 3; }
__host__ __device__ inline int pv_class_cg(int c) { return (c >> 1) / 3; }
// collapsed group of a read from its summary byte (pb2_read_batch.collapsed): ReadExtentions.GetReadCollapsedType (Read.cs:17-64)
__host__ __device__ inline int pv_collapsed_group(int cbyte) {
    if (!(cbyte & 1)) return 0;
    if (cbyte & 2) return 1;
    const int pd = (cbyte >> 2) & 3;
    return pd == 1 ? 2 : (pd == 2 ? 3 : 0);
}
// ReadCollapsedType + 1 of (collapsed group, direction), 0 = none
__host__ __device__ inline int pv_collapsed_code(int cg, int dir) {
    if (cg == 1) return (dir == 2 ? 0 : 1) + 1;   // DuplexStitched / DuplexNonStitched
    if (cg == 2) return (dir == 2 ? 4 : 5) + 1;   // SimplexForward(Non)Stitched
    if (cg == 3) return (dir == 2 ? 6 : 7) + 1;   // SimplexReverse(Non)Stitched
    return 0;
}

struct PvertPileup {
    const uint8_t* data;        // blocks of 1 KB
    const int2* row_meta;       // [n_rows]: (Read.Position, Read.EndPosition) of the row's read; deletion rows: (INT32_MIN + anchor bin, 0)
    const int64_t* tile_row0;   // [n_tiles + 1] first row of a tile (multiple of 32)
    const int32_t* cls_end;     // [n_tiles][n_classes] rows of the tile up to and including this class (each class padded to a multiple of 32)
    int32_t n_classes;
    const uint8_t* ref_base;    // [n_loci] ASCII
    const int32_t* positions;   // [n_loci] or nullptr
    int32_t first_position;
    int64_t n_loci;
    int32_t n_tiles;
};

// 8 x 8 bit-matrix transpose of the eight bytes of x: byte k of the result holds bit k of byte 0..7 of x in its bits 0..7.
__host__ __device__ inline unsigned long long pv_transpose8x8(unsigned long long x) {
    unsigned long long t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull; x = x ^ t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x = x ^ t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x = x ^ t ^ (t << 28);
    return x;
}

struct ReadsView;
struct RegionView;
// reads -> rows: count rows per (tile, class); layout (pad, prefix); fill (rows + row_meta + flagged-entry side list); bit transposition in place
// complex: [1 + n_reads] ints, complex[0] = 0 before the count pass, which lists there the reads that need the general walker (fill: n_complex of them)
cudaError_t launch_pvert_count(const ReadsView& rv, const RegionView& rg, const int32_t* end_pos, int n_classes, int32_t* cls_rows, int32_t* complex, cudaStream_t st);
cudaError_t launch_pvert_layout(int32_t* cls_rows /* in: rows per class; out: inclusive padded prefix */, int32_t n_tiles, int n_classes, int64_t* tile_rows, cudaStream_t st);
cudaError_t launch_pvert_fill(const ReadsView& rv, const RegionView& rg, const int32_t* end_pos, int n_classes, const int64_t* tile_row0, const int32_t* cls_end, int32_t* cursor, uint8_t* data,
                              int2* row_meta, int32_t* row_amp /* optional: amplicon id per base row */, uint32_t* exc_entries, unsigned long long* exc_count, int64_t exc_capacity,
                              const uint8_t* ref_slot, int32_t* complex, int64_t n_complex, cudaStream_t st);
// AmpliconBiasCalculator.Compute over a stream of records (n_dev: device counter, else n_host; valid: optional per-record flag byte, bit 0): sets the
// AmpliconBias filter bit of the SNV records it detects bias for; *status = 1 when a called SNV's position holds more than six amplicon names
cudaError_t launch_pvert_amplicon_bias(const PvertPileup& in, const int32_t* row_amp, pb2_call_record* records, const unsigned long long* n_dev, int64_t n_host, int64_t capacity,
                                       const uint8_t* valid, float acceptance, int min_bq, int* status, int num_sms, cudaStream_t st);
cudaError_t launch_pvert_transpose(uint8_t* data, int64_t n_blocks, cudaStream_t st);
// ref_base[n_loci] ASCII; ref_slot (optional, [n_loci rounded up to 32]): allele2 << 6 of the reference base, 1 where it is not A/C/G/T
cudaError_t launch_pvert_ref_bases(const uint8_t* chr, int64_t chr_len, const int32_t* positions, int32_t first_position, int64_t n_loci, uint8_t* ref_base, uint8_t* ref_slot,
                                   cudaStream_t st);
// 198-bin counts (+ collapsed-read counts) of requested loci straight from the blocks: same outputs as launch_gather_locus_counts
cudaError_t launch_pvert_gather(const PvertPileup& in, const int32_t* req_locus, int32_t n_req, int32_t* out_counts, int32_t* out_collapsed, int min_bq, cudaStream_t st);
// the hot kernel over PVERT (pb2_kernels.cu)
cudaError_t launch_pvert_hot_kernel(const PvertPileup& in, const HotInputsExtra& ex, const HotOutputs& out, const DeviceConfig& cfg, int num_sms, int* tile_counter,
                                    cudaStream_t stream);

}  // namespace pb2
