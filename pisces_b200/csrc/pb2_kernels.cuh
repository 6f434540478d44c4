// Kernels of the per-locus hot path (sm_100a). See DESIGN.md for the data layout and the roofline of each kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pisces_b200.h"

namespace pb2 {

constexpr int kNumAlleles = 6;
constexpr int kNumDirs = 3;
constexpr int kAnchorK = 5;                 // TrackedAnchorSize
constexpr int kNumAnchors = 2 * kAnchorK + 1;
constexpr int kNumBins = kNumAlleles * kNumDirs * kNumAnchors;  // 198 = RegionState._alleleCounts[pos,...]
constexpr int kNumCollapsed = 8;
constexpr int kTileLoci = 32;               // loci per tile = lanes per warp
constexpr int kChunk = 16;                  // entries per lane per step = one 16-byte load per plane
constexpr int kHotThreads = 512;            // one persistent CTA per SM, thread-private 16-bit histograms in shared memory
constexpr int kNarrowThreads = 1024;        // 8-bit histograms: twice the warps in the same shared memory (depth per locus < kNarrowMaxDepth)
constexpr int kNarrowMaxDepth = 60000;      // 8-bit row-wrap counters hold 255 wraps of 256

// Scalars the kernels need from pb2_config (validated / derived on the host, Config semantics of VariantCallingParameters.Validate).
// HotInputsExtra.gapped_ref[locus]: RegionState._gappedMnvReferenceCounts in the low bits; this bit = do not derive SNV candidates from the counts at
// this locus (its SNV candidates are explicit, pb2_explicit.cu:explicit_materialize_snvs)
constexpr int32_t kSuppressCountSnvs = 1 << 30;

struct DeviceConfig {
    int min_bq;
    int noise_level;
    int noise_model;
    float min_frequency, min_frequency_filter, target_lod, variant_freq_filter;
    int max_vq, min_vq, vq_filter, max_gq, min_gq, low_gq_filter, min_coverage, low_depth_filter;
    int rmxn_max_len, rmxn_min_reps;
    float rmxn_freq_limit;
    float sb_acceptance;
    int sb_model, filter_single_strand;
    float no_call_filter;
    int output_gvcf, expect_stitched, expect_collapsed, have_intervals, want_qsum;
    uint32_t one;           // always 1 (see rows_of_word)
    int tune_prefetch;      // tuning experiments: 0 default (2 steps ahead), 1 none, 3, 4
    int tune_ctas_per_sm;   // resident CTAs per SM of the hot kernel (launch bound -> register budget): 4 (64 registers) or 3 (85)
    int snv_from_counts;    // 1: SNV candidates = the counts (CallMNVs off); 0: SNVs are explicit candidates from the finder's state machine
    int ploidy;             // PloidyModel of THIS chromosome (GenotypeCreator.GetPloidyForThisChr): 0 Somatic, 1 DiploidByThresholding, 3 Haploid
    float diploid_minor_vf, diploid_major_vf, diploid_sum_vf;   // DiploidSNVThresholdingParameters
    double sb_min_vf;       // (double)_config.MinFrequency: minDetectableSNP of the Diploid strand-bias model (StrandBiasCalculator.cs:137-148)
    int own_lo, own_hi;     // interval shard: only positions in [own_lo, own_hi] are emitted (pb2_set_owned_range); own_hi = 0: everything
    double vq_error_rate;   // MathOperations.QtoP(noise_level) (VariantQualityCalculator.cs:31)
    double sb_noise;        // Math.Pow(10, -1*noise_level/10f) (StrandBiasCalculator.cs:32)
};

// Device-resident, tile-interleaved pileup ("PTILE32").
struct TilePileup {
    const uint8_t* cq;          // code + quality plane: per tile step the 16-byte code chunks of the active lanes, then their quality chunks
    const uint8_t* anch;        // anchor / collapsed-type plane (same chunk order, one chunk run per step)
    const int64_t* tile_base;   // [n_tiles] first byte of the tile in each plane (multiple of 16)
    const int32_t* depth;       // [n_loci] entries in the source pileup
    const int32_t* pad;         // [n_loci] PAD entries the staged locus carries (chunk tail + dropped low-quality deletions)
    const uint8_t* ref_base;    // [n_loci] ASCII
    const int32_t* positions;   // [n_loci] or nullptr
    int32_t first_position;
    int64_t n_loci;
    int32_t n_tiles;
    int64_t plane_bytes;        // size of the anchor plane (multiple of 16, >= 16); the code + quality plane is twice that
    // "PNIB16" (DESIGN.md 3): the same pileup split by direction and nibble-packed, read by pileup_nib_score_kernel when present (nib != nullptr).
    // Sub-locus (locus, direction in {F, R}) = one lane; sub-tile = 16 loci = 32 sub-loci; per step the 16-byte quality chunks of the active sub-loci,
    // then their 8-byte code chunks (16 allele nibbles), the step padded to 16 bytes.
    const uint8_t* nib;
    const int64_t* nib_tile_base;   // [n_nib_tiles + 1] byte offset of a sub-tile (last entry: total bytes)
    const int32_t* nib_store;       // [2 * n_loci] entries stored for (locus, direction): A/C/G/T of any quality and countable deletions
    const int32_t* nib_depth;       // [2 * n_loci] entries counted for (locus, direction): the stored ones + N bases
    int32_t n_nib_tiles;
    int32_t nib_max_store;
};
constexpr int kNibLoci = 16;        // loci per PNIB16 sub-tile
constexpr int kNibMaxChunks = 256;  // chunks per sub-locus the PNIB16 staging handles (deeper loci stay with the PTILE32 kernel)

// A locus whose SNV candidates passed the cheap callability bars: scored by score_pending_kernel. 96 bytes.
struct PendingLocus {
    int32_t locus;
    int32_t cand_mask;   // bit a: allele a is a candidate; 0x100: a non-point variant was called here; 0x200: any count at the locus
    int32_t c[18];       // [allele][direction] anchor-summed counts
    int32_t gapped;
    int32_t pad_;
    double qsum;
};
static_assert(sizeof(PendingLocus) == 96, "PendingLocus layout");

struct HotOutputs {
    pb2_call_record* ref_records;   // [n_loci] dense reference stream (gVCF) or nullptr
    uint8_t* ref_valid;             // [n_loci]
    pb2_call_record* var_records;   // compacted variant stream
    unsigned long long* var_count;
    int64_t var_capacity;
    PendingLocus* pending;          // queue of loci for score_pending_kernel
    unsigned long long* pending_count;
    int64_t pending_capacity;
    uint32_t* exc_entries;          // flagged mismatching entries: {locus_index, code | qual<<8 | anchor<<16}
    unsigned long long* exc_count;
    int64_t exc_capacity;
    int32_t* counts_out;            // optional [n_loci][198] parity dump (pb2_get_counts); nullptr on the hot path
    int32_t* collapsed_out;         // optional [n_loci][8]
};

struct HotInputsExtra {
    const int32_t* gapped_ref;      // [n_loci] or nullptr (RegionState._gappedMnvReferenceCounts)
    const uint8_t* locus_has_variant;  // [n_loci] or nullptr: a non-point variant was called at this position (ref pruning, AlleleCaller.cs:146-147)
    const uint8_t* chr_seq;         // upper-case chromosome (RMxN), or nullptr
    int64_t chr_len;
    const double* q_to_p_table;     // QtoP(q), q = 0..q_table_max (MathOperations.cs:7-10), or nullptr
    int q_table_max;
    const double* gq_tail_table;    // [kGqTailMaxCov][kGqTailMaxA] Poisson tails of the somatic GQ (pb2_math.cuh:somatic_gq), or nullptr; followed by
                                    // int32 [kGqTailMaxCov][kGqTailMaxA]: the finished GQ when the variant q-score is gq_capped_vq
    int gq_capped_vq;
};

// fills the somatic-GQ tail table for this target LOD (once per handle)
cudaError_t launch_gq_tail_fill(double* table, float target_lod, double p1_capped, int min_gq, int max_gq, cudaStream_t stream);
size_t hot_kernel_smem_bytes(bool narrow, bool collapsed);
cudaError_t launch_hot_kernel(const TilePileup& in, const HotInputsExtra& ex, const HotOutputs& out, const DeviceConfig& cfg, int num_sms, int* tile_counter,
                              int max_depth, cudaStream_t stream);

// CSR -> PTILE32 staging
// PB2_LAYOUT_PACKED2 -> the three planes the scatter reads (in place for code / qual, anchor plane written), then the sparse candidate flags
cudaError_t launch_unpack_packed2(uint8_t* code, uint8_t* qual, uint8_t* anch, int64_t n, cudaStream_t stream);
cudaError_t launch_apply_entry_flags(uint8_t* code, const int64_t* flag_index, const uint8_t* flag_bits, int64_t n_flags, int64_t e0, int64_t e1, cudaStream_t stream);
cudaError_t launch_tile_layout(const int64_t* csr_offsets, int64_t n_loci, int32_t* depth, int64_t* tile_chunks /*[n_tiles]*/, int32_t* max_depth,
                               cudaStream_t stream);
// tiles [tile0, tile0 + n_tiles); code/qual/anch point at entry `entry_base` of the CSR planes (chunked staging)
cudaError_t launch_tile_scatter(const int64_t* csr_offsets, const uint8_t* code, const uint8_t* qual, const uint8_t* anch, int64_t n_loci, int32_t tile0,
                                int32_t n_tiles, int64_t entry_base, const int64_t* tile_base, const uint8_t* ref_base, int min_bq, uint8_t* tcq, uint8_t* tanch, int32_t* pad,
                                uint32_t* exc_entries, unsigned long long* exc_count, int64_t exc_capacity,
                                cudaStream_t stream);
// PTILE32 -> PNIB16 staging: count (per sub-locus sizes, per sub-tile bytes; flags[0] = a Stitched-direction entry exists, flags[1] = max stored), then scatter
cudaError_t launch_nib_count(const TilePileup& in, int32_t* nib_store, int32_t* nib_depth, int64_t* nib_tile_bytes, int32_t* flags, cudaStream_t stream);
cudaError_t launch_nib_scatter(const TilePileup& in, const int32_t* nib_store, const int64_t* nib_tile_base, uint8_t* nib, cudaStream_t stream);
cudaError_t exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, void* temp, size_t temp_bytes, size_t* temp_needed, cudaStream_t stream);

}  // namespace pb2
