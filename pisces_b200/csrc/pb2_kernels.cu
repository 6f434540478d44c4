// Hot-path kernels for sm_100a: tile staging and the fused pileup-count + score kernel.
//
// Replaces (reference @ /root/reference):
//   RegionStateManager.AddAlleleCounts / RegionState.AddAlleleCount     src/lib/Pisces.Processing/RegionState/RegionStateManager.cs:118-220, RegionState.cs:225-239
//   RegionState.GetAllCandidates (reference candidates)                 RegionState.cs:383-453
//   CoverageCalculator.CalculateSinglePoint                             src/lib/Pisces.Calculators/CoverageCalculator.cs:49-98
//   AlleleCaller.ProcessVariant / IsCallable / ComputeGenotypeAndFilterAllele   src/exe/Pisces/Logic/VariantCalling/AlleleCaller.cs:143-177,208-258
//   AlleleProcessor.ApplyFilters, RMxNCalculator                        AlleleProcessor.cs:25-71, src/lib/Pisces.Calculators/RMxNCalculator.cs:19-133
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "pb2_kernels.cuh"
#include "pb2_math.cuh"
#include "pb2_pvert.cuh"

namespace pb2 {

constexpr uint32_t kStagedN = 7;                 // staged allele code of N (the external pb2_pileup_csr code is AlleleType.N = 4)
constexpr uint32_t kPadCode4 = 0x07070707u;      // four PAD code bytes: N | Forward
constexpr uint32_t kQualMask = 0x7fu;            // staged quality bytes carry bit 7 (see tile_scatter_kernel)

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ uint4 ldg_stream(const uint8_t* p) {  // 16-byte streaming load: read once, do not pollute L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void prefetch_l2(const uint8_t* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// branch-free both ways (a switch here compiles to divergent branches in the per-locus tail of the hot kernel)
__device__ __forceinline__ int allele_of_base(uint8_t c) {  // AlleleHelper.GetAlleleType (Utility/AlleleHelper.cs:13-32)
    int a = AT_N;
    a = c == 'A' ? AT_A : a;
    a = c == 'C' ? AT_C : a;
    a = c == 'G' ? AT_G : a;
    a = c == 'T' ? AT_T : a;
    return a;
}
__device__ __forceinline__ char base_of_allele(int a) {   // AT_A 0, AT_G 1, AT_C 2, AT_T 3, everything else 'N'
    static_assert(AT_A == 0 && AT_G == 1 && AT_C == 2 && AT_T == 3, "allele codes");
    const unsigned lut = 'A' | ('G' << 8) | ('C' << 16) | ((unsigned)'T' << 24);
    return (unsigned)a < 4u ? (char)((lut >> (8 * a)) & 0xffu) : 'N';
}

// ------------------------------------------------------------------------------------------------ CSR -> PTILE32
// depth[i] = off[i+1]-off[i];  tile_chunks[t] = bytes per plane of tile t = 16 * sum over the tile's loci of ceil(depth/16)
__global__ void tile_layout_kernel(const int64_t* __restrict__ off, int64_t n_loci, int32_t* __restrict__ depth, int64_t* __restrict__ tile_chunks,
                                   int32_t* __restrict__ max_depth) {
    const int64_t locus = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // blockDim multiple of 32, tiles are warp-aligned
    int d = 0;
    if (locus < n_loci) { d = (int)(off[locus + 1] - off[locus]); depth[locus] = d; }
    int chunks = (d + kChunk - 1) / kChunk;
    const int dmax = __reduce_max_sync(0xffffffffu, d);
    if ((threadIdx.x & 31) == 0 && dmax > 0) atomicMax(max_depth, dmax);
    chunks = __reduce_add_sync(0xffffffffu, chunks);
    if ((threadIdx.x & 31) == 0 && (locus / kTileLoci) * (int64_t)kTileLoci < n_loci) tile_chunks[locus / kTileLoci] = (int64_t)chunks * kChunk;
}

// one warp per tile: lane i copies locus i's entries chunk by chunk into the interleaved position. Entries are normalised on the way so
// that the hot loop is branch-free:
//   * a Deletion entry below the quality bar is never counted (RegionStateManager.cs:170-177) -> replaced by a PAD entry;
//   * the tail of a locus' last 16-entry chunk is filled with PAD entries;
//   * PAD = (code N|Forward, qual 255, anchor 0): it lands in bin [N][Forward][0] and pad[i] of them are subtracted at read-out;
//   * N is staged as allele code 7 (kStagedN) so that the quality rule `q < minBQ -> N` is a byte-parallel OR with 7 in the hot loop;
//   * the quality byte is staged with bit 7 set (q | 0x80, q clamped to 127: Phred qualities end at 93) so that the hot loop's byte-parallel
//     `q - minBQ` cannot borrow across bytes and its sign test is one byte-permute; readers of the plane mask with 0x7f (kQualMask);
//   * candidate flags are kept only where they can matter: the base is a usable (q >= minBQ, A/C/G/T) mismatch against an A/C/G/T
//     reference base. Every flagged entry the hot kernel meets is then a real SNV-candidate exception.
__global__ void tile_scatter_kernel(const int64_t* __restrict__ off, const uint8_t* __restrict__ code, const uint8_t* __restrict__ qual,
                                    const uint8_t* __restrict__ anch, int64_t n_loci, int32_t tile0, int32_t n_tiles, int64_t entry_base,
                                    const int64_t* __restrict__ tile_base, const uint8_t* __restrict__ ref_base, int min_bq, uint8_t* __restrict__ tcq,
                                    uint8_t* __restrict__ tanch, int32_t* __restrict__ pad, uint32_t* __restrict__ exc_entries, unsigned long long* __restrict__ exc_count,
                                    int64_t exc_capacity) {
    const int lane = threadIdx.x & 31;
    const int64_t wi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wi >= n_tiles) return;
    const int64_t tile = tile0 + wi;
    const int64_t locus = tile * kTileLoci + lane;
    if (tile * kTileLoci >= n_loci) return;
    int64_t src = 0;
    int d = 0;
    int ref_allele = AT_N;
    if (locus < n_loci) { src = off[locus]; d = (int)(off[locus + 1] - src); src -= entry_base; ref_allele = allele_of_base(ref_base[locus]); }
    const int nchunks = (d + kChunk - 1) / kChunk;
    const int max_chunks = __reduce_max_sync(0xffffffffu, nchunks);
    int64_t base = tile_base[tile];
    int npad = nchunks * kChunk - d;
    for (int c = 0; c < max_chunks; c++) {
        const bool active = c < nchunks;
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (active) {
            const int slot = __popc(m & ((1u << lane) - 1));
            const int64_t dst = base + (int64_t)slot * kChunk;
            const int n = min(kChunk, d - c * kChunk);
            const int64_t s = src + (int64_t)c * kChunk;
            uint32_t wc[4] = {kPadCode4, kPadCode4, kPadCode4, kPadCode4};
            uint32_t wq[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
            uint32_t wa[4] = {0u, 0u, 0u, 0u};
            for (int k = 0; k < n; k++) {
                uint32_t cb = code[s + k], qb = qual[s + k], ab = anch[s + k];
                const int allele = cb & 7;
                const bool lowq = (int)qb < min_bq;
                if (allele > AT_DEL || ((cb >> 3) & 3) > DIR_S || (ab & 15) >= kNumAnchors || (allele == AT_DEL && lowq)) {
                    cb = kStagedN; qb = 0xff; ab = 0; npad++;       // not countable: PAD
                } else if (lowq || allele >= AT_N || ref_allele == AT_N || allele == ref_allele) {
                    cb &= 0x1f;                                     // flags cannot matter here
                }
                if (cb & 0xe0u) {
                    // a flagged usable mismatch: SNV-candidate bookkeeping the counts cannot express (open ends, '='/'X' support). It is a property of
                    // the entry, not of the counting, so it goes to the segment's side list here and the hot loop never looks at flag bits.
                    const unsigned long long slot = atomicAdd(exc_count, 1ull);
                    if ((int64_t)slot < exc_capacity) { exc_entries[2 * slot] = (uint32_t)locus; exc_entries[2 * slot + 1] = cb | (qb << 8) | (ab << 16); }
                    cb &= 0x1f;
                }
                if ((cb & 7) == AT_N) cb |= kStagedN;
                qb = min(qb, 127u) | 0x80u;
                const int sh = (k & 3) * 8;
                wc[k >> 2] = (wc[k >> 2] & ~(0xffu << sh)) | (cb << sh);
                wq[k >> 2] = (wq[k >> 2] & ~(0xffu << sh)) | (qb << sh);
                wa[k >> 2] = (wa[k >> 2] & ~(0xffu << sh)) | (ab << sh);
            }
            // code and quality share one plane: per step the code chunks of the active lanes, then their quality chunks
            const int64_t dcq = 2 * base + (int64_t)slot * kChunk;
            *reinterpret_cast<uint4*>(tcq + dcq) = make_uint4(wc[0], wc[1], wc[2], wc[3]);
            *reinterpret_cast<uint4*>(tcq + dcq + (int64_t)__popc(m) * kChunk) = make_uint4(wq[0], wq[1], wq[2], wq[3]);
            *reinterpret_cast<uint4*>(tanch + dst) = make_uint4(wa[0], wa[1], wa[2], wa[3]);
        }
        base += (int64_t)__popc(m) * kChunk;
    }
    if (locus < n_loci) pad[locus] = npad;
}

// ------------------------------------------------------------------------------------------------ PB2_LAYOUT_PACKED2 (host pushes of 2 B / entry)
__global__ void unpack_packed2_kernel(uint8_t* __restrict__ code, uint8_t* __restrict__ qual, uint8_t* __restrict__ anch, int64_t n) {
    // 16 entries per thread, 16-byte accesses (the staging buffers are 256-byte aligned); the tail is handled byte by byte
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i >= n) return;
    if (i + 16 <= n) {
        uint4 c = *reinterpret_cast<const uint4*>(code + i), q = *reinterpret_cast<const uint4*>(qual + i), a;
        const uint32_t* cw = &c.x; const uint32_t* qw = &q.x; uint32_t* aw = &a.x;
        uint32_t co[4], qo[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            aw[k] = ((cw[k] >> 5) & 0x07070707u) | ((qw[k] >> 4) & 0x08080808u);
            co[k] = cw[k] & 0x1f1f1f1fu;
            qo[k] = qw[k] & 0x7f7f7f7fu;
        }
        *reinterpret_cast<uint4*>(code + i) = make_uint4(co[0], co[1], co[2], co[3]);
        *reinterpret_cast<uint4*>(qual + i) = make_uint4(qo[0], qo[1], qo[2], qo[3]);
        *reinterpret_cast<uint4*>(anch + i) = a;
    } else {
        for (int64_t k = i; k < n; k++) {
            const uint8_t c = code[k], q = qual[k];
            anch[k] = (uint8_t)((c >> 5) | ((q >> 7) << 3));
            code[k] = c & 0x1f;
            qual[k] = q & 0x7f;
        }
    }
}
__global__ void apply_entry_flags_kernel(uint8_t* __restrict__ code, const int64_t* __restrict__ flag_index, const uint8_t* __restrict__ flag_bits, int64_t n_flags,
                                         int64_t e0, int64_t e1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_flags) return;
    const int64_t e = flag_index[i];
    if (e >= e0 && e < e1) code[e - e0] |= (uint8_t)(flag_bits[i] & 0xe0);
}
cudaError_t launch_unpack_packed2(uint8_t* code, uint8_t* qual, uint8_t* anch, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int64_t threads = (n + 15) / 16;
    unpack_packed2_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(code, qual, anch, n);
    return cudaGetLastError();
}
cudaError_t launch_apply_entry_flags(uint8_t* code, const int64_t* flag_index, const uint8_t* flag_bits, int64_t n_flags, int64_t e0, int64_t e1, cudaStream_t stream) {
    if (n_flags <= 0 || e1 <= e0) return cudaSuccess;
    apply_entry_flags_kernel<<<(unsigned)((n_flags + 255) / 256), 256, 0, stream>>>(code, flag_index, flag_bits, n_flags, e0, e1);
    return cudaGetLastError();
}

cudaError_t launch_tile_layout(const int64_t* off, int64_t n_loci, int32_t* depth, int64_t* tile_chunks, int32_t* max_depth, cudaStream_t stream) {
    const int threads = 256;
    const int64_t n_pad = (n_loci + kTileLoci - 1) / kTileLoci * kTileLoci;
    const unsigned blocks = (unsigned)((n_pad + threads - 1) / threads);
    if (blocks) tile_layout_kernel<<<blocks, threads, 0, stream>>>(off, n_loci, depth, tile_chunks, max_depth);
    return cudaGetLastError();
}
cudaError_t launch_tile_scatter(const int64_t* off, const uint8_t* code, const uint8_t* qual, const uint8_t* anch, int64_t n_loci, int32_t tile0, int32_t n_tiles,
                                int64_t entry_base, const int64_t* tile_base, const uint8_t* ref_base, int min_bq, uint8_t* tcq, uint8_t* tanch, int32_t* pad,
                                uint32_t* exc_entries, unsigned long long* exc_count, int64_t exc_capacity, cudaStream_t stream) {
    const int threads = 256;
    const unsigned blocks = (unsigned)(((int64_t)n_tiles * 32 + threads - 1) / threads);
    if (blocks) tile_scatter_kernel<<<blocks, threads, 0, stream>>>(off, code, qual, anch, n_loci, tile0, n_tiles, entry_base, tile_base, ref_base, min_bq, tcq, tanch, pad, exc_entries, exc_count,
                                                                 exc_capacity);
    return cudaGetLastError();
}

}  // namespace pb2
// (key, index) pairs of the job sink by key (pb2_sink_sort): CUB radix sort over the 40 key bits in use
cudaError_t sink_sort_pairs(void* temp, size_t& temp_bytes, const unsigned long long* kin, unsigned long long* kout, const uint32_t* vin, uint32_t* vout, int64_t n,
                            cudaStream_t st) {
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kin, kout, vin, vout, (int)n, 0, 48, st);
}
namespace pb2 {
// out[i] = sum_{j<i} in[j]   (staging step, not on the hot path: CUB device scan). tile_layout_kernel already scaled the input to bytes.
cudaError_t exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, void* temp, size_t temp_bytes, size_t* temp_needed, cudaStream_t stream) {
    size_t need = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n, stream);
    if (e != cudaSuccess) return e;
    if (temp_needed) *temp_needed = need;
    if (temp == nullptr) return cudaSuccess;
    if (temp_bytes < need) return cudaErrorInvalidValue;
    return cub::DeviceScan::ExclusiveSum(temp, need, in, out, (int)n, stream);
}

// ------------------------------------------------------------------------------------------------ the fused hot kernel
// Thread-private histogram cell for (bin, thread): 16-bit counters interleaved so that the 32 lanes of a warp always hit 32 different
// banks whatever bins they address: halfword index = bin * kHotThreads + (warp >> 1) * 64 + lane * 2 + (warp & 1).
template <typename Cnt>
__device__ __forceinline__ int hist_slot(int warp, int lane) {
    constexpr int kPerWord = 4 / (int)sizeof(Cnt);   // counters per 32-bit word: the warps sharing a word differ, the 32 lanes hit 32 banks
    return (warp / kPerWord) * (32 * kPerWord) + lane * kPerWord + (warp % kPerWord);
}

// RMxNCalculator.ComputeRMxNLengthForIndel (:49-95) on the device-resident chromosome; variant_bases has length <= 1 for point alleles
// `cap`: counting stops at cap repeats; callers only compare min/max of these counts against cap, which capping preserves.
__device__ int rmxn_length_for_indel(int variant_position, const char* vb, int length, const uint8_t* __restrict__ ref, int64_t ref_len, int max_unit, int cap) {
    int best = 0;
    const int first = length - min(max_unit, length);
    for (int pass = 0; pass < 2; pass++) {          // prefixes, then suffixes (bookends)
        for (int i = first; i < length; i++) {
            const int blen = length - i;
            const char* book = pass == 0 ? vb : vb + i;
            int64_t back = variant_position;
            for (int steps = 0; steps < cap; steps++) {   // going back further than cap units cannot change a count that is capped at cap
                const int64_t nb = back - blen;
                if (nb < 0) break;
                bool eq = true;
                for (int k = 0; k < blen; k++) if (ref[nb + k] != (uint8_t)book[k]) { eq = false; break; }
                if (!eq) break;
                back = nb;
            }
            int reps = 0;
            int64_t cur = back;
            while (true) {
                if (cur + blen > ref_len) break;
                bool eq = true;
                for (int k = 0; k < blen; k++) if (ref[cur + k] != (uint8_t)book[k]) { eq = false; break; }
                if (!eq) break;
                reps++;
                cur += blen;
                if (reps >= cap) break;
            }
            best = max(best, reps);
        }
    }
    return best;
}
// RMxNCalculator.ShouldFilter for an SNV (:19-38,104-133)
__device__ bool rmxn_should_filter_snv(int position, char ref_base, char alt_base, float freq, const DeviceConfig& cfg, const uint8_t* __restrict__ ref, int64_t ref_len) {
    if (freq >= cfg.rmxn_freq_limit) return false;
    if (cfg.rmxn_max_len < 0 || cfg.rmxn_min_reps < 0 || ref == nullptr) return false;
    const int cap = max(cfg.rmxn_min_reps, 1);
    const int c1 = rmxn_length_for_indel(position - 1, &ref_base, 1, ref, ref_len, cfg.rmxn_max_len, cap);
    const int i1 = rmxn_length_for_indel(position + 1 - 1, &alt_base, 1, ref, ref_len, cfg.rmxn_max_len, cap);
    const int i2 = rmxn_length_for_indel(position - 1, &alt_base, 1, ref, ref_len, cfg.rmxn_max_len, cap);
    return min(c1, max(i1, i2)) >= cfg.rmxn_min_reps;
}

struct LocusCounts {
    int c[kNumAlleles][kNumDirs];  // anchor-summed counts  (IAlleleSource.GetAlleleCount defaults: all 11 bins)
    double qsum;                   // Σ over coverage-contributing alleles of the base-quality probabilities
};

// Fill one record for a point allele (Reference or Snv) exactly as ProcessVariant + SetGenotypes would. Returns false when a non-reference
// allele is not callable (AlleleCaller.IsCallable) so nothing is emitted.
// counts of a run-time allele out of an array that must stay in registers (no dynamic indexing: a select chain over the six rows)
__device__ __forceinline__ int count_of(const int (&c)[kNumAlleles][kNumDirs], int allele, int d) {
    int v = 0;
#pragma unroll
    for (int a = 0; a < kNumAlleles; a++) v = (a == allele) ? c[a][d] : v;
    return v;
}

// One point allele (SNV or Reference) of a locus: coverage, q-score, strand bias, filters, genotype, record. kRefOnly: the allele is known at compile
// time to be the locus' reference allele (the per-locus Reference candidate of gVCF mode) - that instance is inlined into the hot kernel, its counts
// and its record living in registers; the general one sits behind the out-of-line score_point_allele below.
// The stages of one point allele, separable so that the queued-locus scorer can spread the independent FP64 chains (the q-score and the three sets of
// strand-bias statistics) over the four lanes of its group: (1) coverage and support out of the counts + the cheap callability bars, (2) q-score and
// strand bias, (3) filters, genotype, record.
struct PointPrep {
    int cov[3], sup[3];
    int total, nocalls, ref_support, allele_support;
    float freq;
    bool is_ref;
};
template <bool kRefOnly>
__device__ __forceinline__ bool point_allele_prepare(const int (&c)[kNumAlleles][kNumDirs], int ref_allele, int alt_allele, int gapped, const DeviceConfig& cfg, PointPrep& p) {
    p.is_ref = kRefOnly || alt_allele == ref_allele;
    p.total = 0; p.nocalls = 0; p.ref_support = 0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        p.cov[d] = c[AT_A][d] + c[AT_C][d] + c[AT_G][d] + c[AT_T][d] + c[AT_DEL][d];
        p.total += p.cov[d];
        p.nocalls += c[AT_N][d];
        p.sup[d] = count_of(c, alt_allele, d);
        if (ref_allele != AT_N) p.ref_support += kRefOnly ? p.sup[d] : count_of(c, ref_allele, d);
    }
    p.allele_support = p.sup[0] + p.sup[1] + p.sup[2];
    if (p.is_ref) p.allele_support = max(0, p.allele_support - gapped);  // CoverageCalculator.cs:94-97
    else p.ref_support = max(0, p.ref_support - gapped);                 // :90-93
    p.freq = allele_frequency(p.allele_support, p.total);
    // cheap callability tests first (same outcome as evaluating Q/SB first, AlleleCaller.cs:236-258)
    if (!p.is_ref) {
        if (p.total < cfg.min_coverage && !cfg.output_gvcf) return false;
        if (p.total != 0 && p.freq < cfg.min_frequency) return false;
    }
    return true;
}
__device__ __forceinline__ int point_allele_vq(const PointPrep& p, double qsum, const DeviceConfig& cfg, int& nl_applied) {
    nl_applied = 0;
    if (p.allele_support <= 0) return 0;
    int nl = cfg.noise_level;
    double error_rate = cfg.vq_error_rate;
    if (cfg.noise_model == 1) {   // NoiseModel.Window (AlleleCaller.cs:215-218)
        nl = (int)(-10 * log10(qsum / p.total));
        error_rate = q_to_p((double)nl);
    }
    nl_applied = nl;
    return (p.total == 0) ? 0 : poisson_qscore(p.allele_support, p.total, error_rate, cfg.max_vq);
}
// RMxNCalculator.ShouldFilter of a point allele (reference alleles are never filtered, AlleleProcessor.cs:44-66): reads the chromosome, nothing else
__device__ __forceinline__ bool point_allele_rmxn(const PointPrep& p, int position, int ref_allele, int alt_allele, const DeviceConfig& cfg, const HotInputsExtra& ex) {
    return !p.is_ref && rmxn_should_filter_snv(position, base_of_allele(ref_allele), base_of_allele(alt_allele), p.freq, cfg, ex.chr_seq, ex.chr_len);
}
__device__ __forceinline__ void point_allele_finish(const PointPrep& p, int vq, int nl_applied, const SbResult& sb, bool rmxn_hit, double qsum, int position, int ref_allele,
                                                    int alt_allele, const DeviceConfig& cfg, const HotInputsExtra& ex, pb2_call_record& r) {
    const bool is_ref = p.is_ref;
    const int total = p.total, nocalls = p.nocalls, allele_support = p.allele_support, ref_support = p.ref_support;
    const float freq = p.freq;
    // AlleleProcessor.Process / ApplyFilters
    const float all_reads = (float)(total + nocalls);
    const float frac_nc = all_reads == 0 ? 0.0f : ((float)nocalls / all_reads);
    unsigned filters = 0;
    if (cfg.low_depth_filter >= 0 && total < cfg.low_depth_filter) filters |= 1u << FLT_LOW_DEPTH;
    if (vq < cfg.vq_filter && total != 0) filters |= 1u << FLT_LOW_VQ;
    if (!is_ref) {
        if (cfg.no_call_filter >= 0 && frac_nc > cfg.no_call_filter) filters |= 1u << FLT_NO_CALL;
        if (!sb.acceptable || (cfg.filter_single_strand && !sb.var_both)) filters |= 1u << FLT_STRAND_BIAS;
        if (rmxn_hit) filters |= 1u << FLT_RMXN;
        if (freq < cfg.variant_freq_filter) filters |= 1u << FLT_LOW_VF;
    }
    // SomaticGenotyper + GQ (per allele). Germline ploidy: exact for a locus whose only allele is this reference allele; the alleles of a locus
    // with variants are genotyped together when pb2_flush merges them (germline_locus_pass)
    const float ref_freq = allele_frequency(ref_support, total);
    int gt, gq;
    if (cfg.ploidy == PLOIDY_SOMATIC) {
        gt = somatic_genotype(is_ref, total, freq, ref_freq, cfg.min_frequency_filter, cfg.min_coverage);
        gq = somatic_gq(gt, vq, total, freq, cfg.target_lod, cfg.min_gq, cfg.max_gq, ex.q_to_p_table, ex.q_table_max, ex.gq_tail_table, ex.gq_capped_vq);
    } else {
        const bool hap = cfg.ploidy == PLOIDY_HAPLOID;
        gt = is_ref ? germline_reference_only_genotype(hap, total, allele_support, ref_support, cfg.diploid_minor_vf, cfg.diploid_major_vf, cfg.min_coverage) : GT_HET_ALT_REF;
        gq = is_ref ? germline_gq(hap, gt, total, allele_support, cfg.min_gq, cfg.max_gq) : 0;
    }
    if (cfg.low_gq_filter >= 0 && (float)gq < (float)cfg.low_gq_filter) filters |= 1u << FLT_LOW_GQ;

    r.position = position;
    r.type = is_ref ? CAT_REF : CAT_SNV;
    r.genotype = (uint8_t)gt;
    r.sb_flags = (sb.acceptable ? 1 : 0) | (sb.var_both ? 2 : 0) | (sb.cov_both ? 4 : 0);
    r.open_flags = 0;
    r.filters = (uint16_t)filters;
    r.noise_level = (uint16_t)nl_applied;
    r.variant_qscore = vq;
    r.genotype_qscore = gq;
    r.total_coverage = total;
#pragma unroll
    for (int d = 0; d < 3; d++) { r.coverage_by_direction[d] = p.cov[d]; r.support_by_direction[d] = p.sup[d]; }
    r.allele_support = allele_support;
    r.reference_support = ref_support;
    r.num_no_calls = nocalls;
    r.fraction_no_calls = frac_nc;
    r.allele_bytes = (uint32_t)(uint8_t)base_of_allele(ref_allele) | ((uint32_t)(uint8_t)base_of_allele(alt_allele) << 8);
    r.ref_len = 1;
    r.alt_len = 1;
    r.sum_base_quality = qsum;
    r.bias_score = sb.bias;
    r.gatk_bias_score = sb.gatk;
}
template <bool kRefOnly>
__device__ __forceinline__ bool score_point_allele_impl(const int (&c)[kNumAlleles][kNumDirs], double qsum, int position, int ref_allele,
                                                        int alt_allele /* == ref_allele for Reference */, int gapped, const DeviceConfig& cfg,
                                                        const HotInputsExtra& ex, pb2_call_record& r) {
    PointPrep p;
    if (!point_allele_prepare<kRefOnly>(c, ref_allele, alt_allele, gapped, cfg, p)) return false;
    int nl_applied;
    const int vq = point_allele_vq(p, qsum, cfg, nl_applied);
    if (!p.is_ref && vq < cfg.min_vq) return false;
    SbResult sb;
    sb.bias = 0; sb.gatk = 0; sb.acceptable = false; sb.var_both = false; sb.cov_both = false;   // new BiasResults()
    if (p.allele_support > 0) sb = strand_bias(p.cov, p.sup, cfg.sb_noise, (double)cfg.sb_acceptance, cfg.sb_model, cfg.sb_min_vf);
    point_allele_finish(p, vq, nl_applied, sb, point_allele_rmxn(p, position, ref_allele, alt_allele, cfg, ex), qsum, position, ref_allele, alt_allele, cfg, ex, r);
    return true;
}

__device__ __noinline__ bool score_point_allele(const LocusCounts& lc, int position, int ref_allele, int alt_allele /* == ref_allele for Reference */, int gapped,
                                   const DeviceConfig& cfg, const HotInputsExtra& ex, pb2_call_record& r) {
    return score_point_allele_impl<false>(lc.c, lc.qsum, position, ref_allele, alt_allele, gapped, cfg, ex, r);
}

__device__ __forceinline__ void store_record(pb2_call_record* dst, const pb2_call_record& r) {
    const uint4* s = reinterpret_cast<const uint4*>(&r);
    uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(pb2_call_record) / 16); i++) d[i] = s[i];
}

// After a locus' counts are known: screen its SNV candidates with the cheap bars of AlleleCaller.IsCallable (:246-251) and queue the locus for
// score_pending_kernel if any survives (the FP64 chain then runs lane-dense instead of at 1/32 lane utilisation); otherwise, in gVCF mode,
// score and emit its reference allele in place.
constexpr int kCtaPending = 64;   // per-CTA queue of the vertical-counter kernel: 64 loci x 4 alleles = one 256-thread scoring pass

// The counts arrive as a plain array indexed with compile-time constants only, so they stay in registers: LocusCounts (whose address the out-of-line
// scorer takes) is built inside the branches that need it, not for every locus (that cost 129 MB of local-memory stores per million loci).
// kRefStream = false: an instance for VCF-only runs (no dense reference stream): the inlined reference-allele scorer is not even compiled in, which is
// what lets the counting loop of the PVERT kernel live in 40 registers (more warps, deeper loads in flight).
template <bool kRefStream = true>
__device__ __forceinline__ void finish_locus(const int (&cnt)[kNumAlleles][kNumDirs], double qsum, int any, int64_t locus, int ref_allele, const TilePileup& in,
                                             const HotInputsExtra& ex, const HotOutputs& out, const DeviceConfig& cfg, PendingLocus* cta_queue = nullptr,
                                             int* cta_count = nullptr, int cta_capacity = kCtaPending) {
    if (cfg.own_hi > 0) {   // an interval shard scores the loci it owns; its halo is staged for the alleles that reach into it, never emitted
        const int position = in.positions ? in.positions[locus] : in.first_position + (int)locus;
        if (position < cfg.own_lo || position > cfg.own_hi) { if (kRefStream && out.ref_records != nullptr) out.ref_valid[locus] = 0; return; }
    }
    const int gapped_word = ex.gapped_ref ? ex.gapped_ref[locus] : 0;
    const int gapped = gapped_word & (kSuppressCountSnvs - 1);
    unsigned cand_mask = 0;
    if (ref_allele != AT_N && cfg.snv_from_counts && !(gapped_word & kSuppressCountSnvs)) {
        int total = 0;
#pragma unroll
        for (int d = 0; d < 3; d++) total += cnt[AT_A][d] + cnt[AT_C][d] + cnt[AT_G][d] + cnt[AT_T][d] + cnt[AT_DEL][d];
#pragma unroll
        for (int alt = 0; alt < 4; alt++) {
            const int sup = cnt[alt][0] + cnt[alt][1] + cnt[alt][2];
            if (alt == ref_allele || sup == 0) continue;
            if (total < cfg.min_coverage && !cfg.output_gvcf) continue;
            if (total != 0 && allele_frequency(sup, total) < cfg.min_frequency) continue;
            cand_mask |= 1u << alt;
        }
    }
    const bool has_ext_variant = ex.locus_has_variant ? (ex.locus_has_variant[locus] != 0) : false;
    if (cand_mask != 0) {
        PendingLocus* dst = nullptr;
        if (cta_queue != nullptr) {
            const int s = atomicAdd(cta_count, 1);
            if (s < cta_capacity) dst = cta_queue + s;
        }
        if (dst == nullptr) {   // no CTA queue (198-bin kernel) or it is full: the global queue, scored by score_pending_kernel
            const unsigned long long slot = atomicAdd(out.pending_count, 1ull);
            if ((int64_t)slot < out.pending_capacity) dst = out.pending + slot;
        }
        if (dst != nullptr) {
            PendingLocus pl;
            pl.locus = (int32_t)locus;
            pl.cand_mask = (int32_t)cand_mask | (has_ext_variant ? 0x100 : 0) | (any > 0 ? 0x200 : 0);
#pragma unroll
            for (int a = 0; a < kNumAlleles; a++)
#pragma unroll
                for (int d = 0; d < kNumDirs; d++) pl.c[a * kNumDirs + d] = cnt[a][d];
            pl.gapped = gapped;
            pl.pad_ = 0;
            pl.qsum = qsum;
            const uint4* sp = reinterpret_cast<const uint4*>(&pl);
            uint4* dp = reinterpret_cast<uint4*>(dst);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(PendingLocus) / 16); i++) dp[i] = sp[i];
        }
        if (kRefStream && out.ref_records != nullptr) out.ref_valid[locus] = 0;   // decided when the queued locus is scored
    } else if (kRefStream && out.ref_records != nullptr) {
        // RegionState.GetAllCandidates: a Reference candidate per position when gVCF; zero-coverage positions only with intervals (:446)
        // ... and never past the chromosome end (:411-412)
        const int position = in.positions ? in.positions[locus] : in.first_position + (int)locus;
        const bool emit = cfg.output_gvcf && !has_ext_variant && (cfg.have_intervals || any > 0) && (ex.chr_len == 0 || position <= ex.chr_len);
        if (emit) {
            pb2_call_record r;   // registers: the inlined scorer writes fields, store_record reads them back as six 16-byte words
            score_point_allele_impl<true>(cnt, qsum, position, ref_allele, ref_allele, gapped, cfg, ex, r);
            store_record(out.ref_records + locus, r);
        }
        out.ref_valid[locus] = emit ? 1 : 0;
    }
}

// Histogram rows are allele-minor: row = allele + 6 * direction, bin = row * 11 + anchor (0..197); [N][Forward][0] (bin 44) also absorbs PADs.
constexpr int kPadBin = (AT_N + 6 * DIR_F) * kNumAnchors + 0;

// Four entries at a time with byte-parallel arithmetic: quality test, N-forcing and the bin index never leave the packed word; only the
// shared-memory read-modify-write is per entry.
__device__ __forceinline__ uint32_t bins_of_word(uint32_t c4, uint32_t q4, uint32_t a4, uint32_t minbq4) {
    uint32_t al4 = c4 & 0x07070707u;
    al4 -= (al4 & (al4 >> 1) & (al4 >> 2) & 0x01010101u) * 3u;        // staged N (7) -> AlleleType.N (4)
    const uint32_t dr6 = ((c4 >> 3) & 0x03030303u) * 6u;            // per byte <= 12
    const uint32_t row4 = al4 + dr6;                                  // <= 17
    const uint32_t rowN4 = dr6 + 0x04040404u;                         // the N row of the same direction
    const uint32_t g7 = (q4 - minbq4) & 0x80808080u;                  // bit 7 of a byte set <=> q >= minBQ (staged q carries bit 7: no borrow crosses bytes)
    const uint32_t m = (g7 - (g7 >> 7)) | g7;                         // 0xFF where q >= minBQ
    const uint32_t rs4 = (row4 & m) | (rowN4 & ~m);                   // RegionStateManager.cs:180-181: low quality -> N
    return rs4 * 11u + (a4 & 0x0f0f0f0fu);                            // per byte <= 197
}

template <typename Cnt, int kThreads, bool kWantQsum, bool kCollapsed>
__device__ __forceinline__ void count_word(uint32_t c4, uint32_t q4, uint32_t a4, uint32_t minbq4, Cnt* __restrict__ my, const double* __restrict__ q_lut,
                                           int min_bq, double& qsum, uint32_t& wrap_acc) {
    const uint32_t bin4 = bins_of_word(c4, q4, a4, minbq4);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t b = (bin4 >> (8 * k)) & 0xffu;
        const uint32_t v = (uint32_t)my[b * kThreads] + 1u;
        my[b * kThreads] = (Cnt)v;
        if (sizeof(Cnt) == 1) wrap_acc |= v;     // bit 8 set <=> an 8-bit counter wrapped somewhere in this chunk
    }
    if (kWantQsum || kCollapsed) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t code = (c4 >> (8 * k)) & 0xffu, q = (q4 >> (8 * k)) & kQualMask, an = (a4 >> (8 * k)) & 0xffu;
            const int allele = code & 7;   // staged: A,G,C,T = 0..3, Del = 5, N = 7
            const bool usable = allele == AT_DEL || (allele < AT_N && (int)q >= min_bq);   // counted as something other than N (PADs are N)
            if (kWantQsum) { if (usable && allele != AT_DEL) qsum += q_lut[q]; }   // Σ over A,C,G,T (Deletion entries carry no base quality, :191)
            if (kCollapsed) {
                const int ct = (int)(an >> 4);
                if (ct != 0 && usable) {   // CollapsedRegionState.AddCollapsedReadCount (:28-44)
                    my[(kNumBins + ct - 1) * kThreads] += 1;
                    if (ct - 1 == 4 || ct - 1 == 6) my[(kNumBins + 2) * kThreads] += 1;
                    else if (ct - 1 == 5 || ct - 1 == 7) my[(kNumBins + 3) * kThreads] += 1;
                }
            }
        }
    }
}

// Rare path of the 8-bit histogram: some counter wrapped while this chunk was added. A bin wrapped iff its value is now smaller than the
// number of entries the chunk put into it; the wrap is recorded per histogram row (allele, direction), which is all the scoring needs.
template <int kThreads>
__device__ __noinline__ void note_wraps(const uint4 wc, const uint4 wq, const uint4 wa, uint32_t minbq4, const uint8_t* __restrict__ my, uint8_t* __restrict__ my_wraps) {
    const uint32_t b4[4] = {bins_of_word(wc.x, wq.x, wa.x, minbq4), bins_of_word(wc.y, wq.y, wa.y, minbq4), bins_of_word(wc.z, wq.z, wa.z, minbq4),
                            bins_of_word(wc.w, wq.w, wa.w, minbq4)};
    for (int k = 0; k < kChunk; k++) {
        const uint32_t b = (b4[k >> 2] >> ((k & 3) * 8)) & 0xffu;
        int cnt = 0;
        bool first = true;
        for (int j = 0; j < kChunk; j++) {
            const uint32_t bj = (b4[j >> 2] >> ((j & 3) * 8)) & 0xffu;
            if (bj == b) { cnt++; if (j < k) first = false; }
        }
        if (first && (int)my[b * kThreads] < cnt) my_wraps[(b / kNumAnchors) * kThreads] += 1;
    }
}

// Cnt = uint16_t, 512 threads: general variant (any depth < 65536 per locus, collapsed-read counts, the full 198-bin dump).
// Cnt = uint8_t, 1024 threads: twice the resident warps for the same shared memory; wraps are caught per chunk and kept per row.
template <typename Cnt, int kThreads, bool kWantQsum, bool kCollapsed>
__global__ void __launch_bounds__(kThreads, 1)
pileup_count_score_kernel(const __grid_constant__ TilePileup in, const __grid_constant__ HotInputsExtra ex, const __grid_constant__ HotOutputs out,
                          const __grid_constant__ DeviceConfig cfg, int* __restrict__ tile_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool kNarrow = sizeof(Cnt) == 1;
    constexpr int kHotThreads = kThreads;
    constexpr int kWrapRows = kNarrow ? kNumAlleles * kNumDirs : 0;
    constexpr int kRows = kNumBins + (kCollapsed ? kNumCollapsed : 0) + kWrapRows;
    Cnt* hist = reinterpret_cast<Cnt*>(smem_raw);                                               // [rows][kThreads]
    double* q_lut = reinterpret_cast<double*>(smem_raw + (size_t)kRows * kThreads * sizeof(Cnt));   // [256] 10^(-q/10f)
    __shared__ int s_tile[kThreads / 32];
    static_assert(!(kNarrow && kCollapsed), "collapsed-read counts use the 16-bit variant");

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    Cnt* my = hist + hist_slot<Cnt>(warp, lane);   // my[bin * kThreads]
    uint8_t* my_wraps = reinterpret_cast<uint8_t*>(my) + (size_t)(kNumBins + (kCollapsed ? kNumCollapsed : 0)) * kThreads;   // [18 rows] (8-bit variant)
    for (int b = 0; b < kRows; b++) my[b * kThreads] = 0;
    if (kWantQsum) {
        for (int q = threadIdx.x; q < 256; q += kThreads) q_lut[q] = pow(10.0, (double)((float)(-q) / 10.0f));  // RegionStateManager.cs:191 (float exponent)
        __syncthreads();
    }
    const uint32_t minbq4 = (uint32_t)cfg.min_bq * 0x01010101u;

    while (true) {
        if (lane == 0) s_tile[warp] = atomicAdd(tile_counter, 1);
        __syncwarp();
        const int tile = s_tile[warp];
        __syncwarp();
        if (tile >= in.n_tiles) break;

        const int64_t locus = (int64_t)tile * kTileLoci + lane;
        const bool have_locus = locus < in.n_loci;
        const int depth = have_locus ? in.depth[locus] : 0;
        const int nchunks = (depth + kChunk - 1) / kChunk;
        const int max_chunks = __reduce_max_sync(0xffffffffu, nchunks);
        const int ref_allele = have_locus ? allele_of_base(in.ref_base[locus]) : AT_N;
        int64_t base = in.tile_base[tile];
        double qsum = 0.0;
        int extra_pad = 0;   // whole PAD chunks counted by lanes that ran out of entries before the longest lane of the tile

        // software pipeline: chunk c+1 is in flight while chunk c is histogrammed; lanes without a chunk process a PAD chunk
        const uint4 pad_c = make_uint4(kPadCode4, kPadCode4, kPadCode4, kPadCode4), pad_q = make_uint4(~0u, ~0u, ~0u, ~0u), pad_a = make_uint4(0, 0, 0, 0);
        uint4 nc = pad_c, nq = pad_q, na = pad_a;
        {
            const bool active = 0 < nchunks;
            const unsigned m = __ballot_sync(0xffffffffu, active);
            if (active) {
                const int64_t o = base + (int64_t)__popc(m & ((1u << lane) - 1)) * kChunk;
                nc = ldg_stream(in.cq + base + o); nq = ldg_stream(in.cq + base + o + (int64_t)__popc(m) * kChunk); na = ldg_stream(in.anch + o);
            }
            base += (int64_t)__popc(m) * kChunk;
        }
        for (int c = 0; c < max_chunks; c++) {
            const uint4 wc = nc, wq = nq, wa = na;
            if (c >= nchunks) extra_pad += kChunk;
            nc = pad_c; nq = pad_q; na = pad_a;
            {
                // chunk c+1 into registers now, chunk c+2 into L2: both requests are in flight for the whole of this iteration
                const bool active = (c + 1) < nchunks;
                const unsigned m = __ballot_sync(0xffffffffu, active);
                if (active) {
                    const int64_t o = base + (int64_t)__popc(m & ((1u << lane) - 1)) * kChunk;
                    nc = ldg_stream(in.cq + base + o); nq = ldg_stream(in.cq + base + o + (int64_t)__popc(m) * kChunk); na = ldg_stream(in.anch + o);
                }
                base += (int64_t)__popc(m) * kChunk;
                const int64_t pf = min(base + lane * kChunk, in.plane_bytes - kChunk);
                prefetch_l2(in.cq + 2 * pf); prefetch_l2(in.cq + 2 * pf + 512); prefetch_l2(in.anch + pf);
                asm volatile("" ::: "memory");   // keep the histogram traffic below the loads
            }
            uint32_t wrap_acc = 0;
            count_word<Cnt, kThreads, kWantQsum, kCollapsed>(wc.x, wq.x, wa.x, minbq4, my, q_lut, cfg.min_bq, qsum, wrap_acc);
            count_word<Cnt, kThreads, kWantQsum, kCollapsed>(wc.y, wq.y, wa.y, minbq4, my, q_lut, cfg.min_bq, qsum, wrap_acc);
            count_word<Cnt, kThreads, kWantQsum, kCollapsed>(wc.z, wq.z, wa.z, minbq4, my, q_lut, cfg.min_bq, qsum, wrap_acc);
            count_word<Cnt, kThreads, kWantQsum, kCollapsed>(wc.w, wq.w, wa.w, minbq4, my, q_lut, cfg.min_bq, qsum, wrap_acc);
            if (kNarrow) { if (wrap_acc & 0x100u) note_wraps<kThreads>(wc, wq, wa, minbq4, reinterpret_cast<const uint8_t*>(my), my_wraps); }
        }

        // ---- read the histogram out (and clear it for the next tile)
        int cnt[kNumAlleles][kNumDirs];
        int any = 0;
        const int npad = (have_locus ? in.pad[locus] : 0) + extra_pad;
#pragma unroll
        for (int a = 0; a < kNumAlleles; a++)
#pragma unroll
            for (int d = 0; d < kNumDirs; d++) {
                int s = 0;
#pragma unroll
                for (int an = 0; an < kNumAnchors; an++) {
                    const int hbin = (a + 6 * d) * kNumAnchors + an;            // shared-memory layout (allele-minor)
                    int v = my[hbin * kHotThreads];
                    my[hbin * kHotThreads] = 0;
                    if (hbin == kPadBin) v -= npad;
                    if (!kNarrow) { if (out.counts_out != nullptr && have_locus) out.counts_out[locus * kNumBins + (a * kNumDirs + d) * kNumAnchors + an] = v; }   // RegionState order
                    s += v;
                }
                if (kNarrow) { s += 256 * (int)my_wraps[(a + 6 * d) * kThreads]; my_wraps[(a + 6 * d) * kThreads] = 0; }
                cnt[a][d] = s;
                any += s;
            }
        if (kCollapsed) {
#pragma unroll
            for (int t = 0; t < kNumCollapsed; t++) {
                const int v = my[(kNumBins + t) * kHotThreads];
                my[(kNumBins + t) * kHotThreads] = 0;
                if (out.collapsed_out != nullptr && have_locus) out.collapsed_out[locus * kNumCollapsed + t] = v;
            }
        }
        if (!have_locus) continue;

        finish_locus(cnt, qsum, any, locus, ref_allele, in, ex, out, cfg);
    }
}

// A queued locus scored by 4 adjacent lanes: lane j takes alternate allele j of (A, C, G, T) (the (ref, alt) order of AlleleCaller.cs:172-176 is
// restored when the records are sorted), then lane 0 the reference allele if nothing was called there (:146-147). Must be called by full warps.
__device__ __noinline__ void score_queued_locus(const PendingLocus* item /* nullptr: idle group */, int j, const TilePileup& in, const HotInputsExtra& ex,
                                                const HotOutputs& out, const DeviceConfig& cfg) {
    // A group of four lanes per queued locus, the whole warp in lockstep. Pass i scores the i-th SNV candidate of every locus of the warp (nearly always
    // the only one); in it the independent FP64 chains run side by side - lane 0 of a group the Poisson q-score, lanes 1..3 the overall / forward /
    // reverse strand-bias statistics (one instruction stream for all three) - and lane 0 finishes the record. The time of a pass is that of the longest
    // chain, not of their sum: it runs after the CTA's last tile, when there is nothing left to hide it behind. Every shuffle is a full-warp one at a
    // warp-uniform point: with per-group masks the groups lose their convergence in the first data-dependent loop and run one after the other.
    const int lane = threadIdx.x & 31;
    const int lead = lane & ~3;
    LocusCounts lc;
#pragma unroll
    for (int a = 0; a < kNumAlleles; a++)
#pragma unroll
        for (int d = 0; d < kNumDirs; d++) lc.c[a][d] = item ? item->c[a * kNumDirs + d] : 0;
    lc.qsum = item ? item->qsum : 0.0;
    const int64_t locus = item ? item->locus : 0;
    const int cand_mask = item ? item->cand_mask : 0, gapped = item ? item->gapped : 0;
    const int ref_allele = item ? allele_of_base(in.ref_base[locus]) : AT_N;
    const int position = in.positions ? in.positions[locus] : in.first_position + (int)locus;
    bool called = false;
    unsigned rem = (unsigned)cand_mask & 0xfu;
    const int passes = __reduce_max_sync(0xffffffffu, __popc(rem));
#pragma unroll 1
    for (int pass = 0; pass < passes; pass++) {
        const int alt = rem ? __ffs((int)rem) - 1 : 0;
        const bool have = rem != 0;
        rem &= rem - 1;
        PointPrep p;
        const bool ok = point_allele_prepare<false>(lc.c, ref_allele, alt, gapped, cfg, p) && have;   // the same on the four lanes of a group
        int vq = 0, nl_applied = 0;
        SbStats st;
        st.fn = st.fp = st.vg = st.coverage = st.support = 0;
        if (ok) {
            if (j == 0) vq = point_allele_vq(p, lc.qsum, cfg, nl_applied);
            else if (p.allele_support > 0) st = sb_stats_of(j - 1, p.cov, p.sup, cfg.sb_noise, cfg.sb_model, cfg.sb_min_vf);
        }
        SbStats o, f, r;
        o.vg = __shfl_sync(0xffffffffu, st.vg, lead + 1); o.fp = __shfl_sync(0xffffffffu, st.fp, lead + 1);
        o.coverage = __shfl_sync(0xffffffffu, st.coverage, lead + 1); o.support = __shfl_sync(0xffffffffu, st.support, lead + 1);
        f.vg = __shfl_sync(0xffffffffu, st.vg, lead + 2); f.fp = __shfl_sync(0xffffffffu, st.fp, lead + 2);
        f.coverage = __shfl_sync(0xffffffffu, st.coverage, lead + 2); f.support = __shfl_sync(0xffffffffu, st.support, lead + 2);
        r.vg = __shfl_sync(0xffffffffu, st.vg, lead + 3); r.fp = __shfl_sync(0xffffffffu, st.fp, lead + 3);
        r.coverage = __shfl_sync(0xffffffffu, st.coverage, lead + 3); r.support = __shfl_sync(0xffffffffu, st.support, lead + 3);
        o.fn = f.fn = r.fn = 0;
        if (ok && j == 0 && vq >= cfg.min_vq) {   // AlleleCaller.IsCallable
            SbResult sb;
            sb.bias = 0; sb.gatk = 0; sb.acceptable = false; sb.var_both = false; sb.cov_both = false;   // new BiasResults()
            if (p.allele_support > 0) sb = strand_bias_combine(o, f, r, (double)cfg.sb_acceptance);
            pb2_call_record rec;
            point_allele_finish(p, vq, nl_applied, sb, point_allele_rmxn(p, position, ref_allele, alt, cfg, ex), lc.qsum, position, ref_allele, alt, cfg, ex, rec);
            called = true;
            const unsigned long long slot = atomicAdd(out.var_count, 1ull);
            if ((int64_t)slot < out.var_capacity) store_record(out.var_records + slot, rec);
        }
        __syncwarp();
    }
    if (item != nullptr && j == 0 && out.ref_records != nullptr) {
        const bool variant_called = ((cand_mask & 0x100) != 0) || called;
        const bool emit = cfg.output_gvcf && !variant_called && (cfg.have_intervals || (cand_mask & 0x200)) && (ex.chr_len == 0 || position <= ex.chr_len);
        if (emit) {
            pb2_call_record r;
            score_point_allele(lc, position, ref_allele, ref_allele, gapped, cfg, ex, r);
            store_record(out.ref_records + locus, r);
        }
        out.ref_valid[locus] = emit ? 1 : 0;
    }
}

// The same queue scored by the whole CTA (256 threads) with one TASK per thread instead of one locus per group of lanes: of a queue of up to 51 loci,
// threads 0..50 take the Poisson q-score of locus t, threads 51..203 one of the three sets of strand-bias statistics of locus (t - 51) / 3, threads
// 204..254 the RMxN scan of the chromosome around locus t - 204. A warp then holds (nearly) one kind of task - no divergent instruction streams to
// serialise - and the independent chains of a locus run on different warps at the same time; the results meet in shared memory, and thread t finishes
// the record of locus t. This pass runs after the CTA's last tile with nothing to hide behind: its time is that of the slowest single chain.
constexpr int kCtaTaskLoci = 51;
struct QueueScratch {
    double vg[kCtaTaskLoci][3], fp[kCtaTaskLoci][3];
    uint8_t rmxn[kCtaTaskLoci];
};
__device__ __noinline__ void score_cta_queue(const PendingLocus* q, int n, QueueScratch& sc, const TilePileup& in, const HotInputsExtra& ex, const HotOutputs& out,
                                             const DeviceConfig& cfg) {
    const int t = threadIdx.x;
    const bool lead = t < kCtaTaskLoci;
    const bool scan = t >= 4 * kCtaTaskLoci && t < 5 * kCtaTaskLoci;
    const int item = lead ? t : (scan ? t - 4 * kCtaTaskLoci : (t - kCtaTaskLoci) / 3);
    const int which = (lead || scan) ? 0 : (t - kCtaTaskLoci) % 3;
    const bool have_item = item < n && t < 5 * kCtaTaskLoci;
    LocusCounts lc;
#pragma unroll
    for (int a = 0; a < kNumAlleles; a++)
#pragma unroll
        for (int d = 0; d < kNumDirs; d++) lc.c[a][d] = have_item ? q[item].c[a * kNumDirs + d] : 0;
    lc.qsum = have_item ? q[item].qsum : 0.0;
    const int64_t locus = have_item ? q[item].locus : 0;
    const int cand_mask = have_item ? q[item].cand_mask : 0, gapped = have_item ? q[item].gapped : 0;
    const int ref_allele = have_item ? allele_of_base(in.ref_base[locus]) : AT_N;
    const int position = in.positions ? in.positions[locus] : in.first_position + (int)locus;
    bool called = false;
    unsigned rem = (unsigned)cand_mask & 0xfu;
    while (__syncthreads_or(rem != 0)) {   // pass i: the i-th SNV candidate of every queued locus (nearly always the only one)
        const int alt = rem ? __ffs((int)rem) - 1 : 0;
        const bool have = rem != 0;
        rem &= rem - 1;
        PointPrep p;
        const bool ok = point_allele_prepare<false>(lc.c, ref_allele, alt, gapped, cfg, p) && have;
        int vq = 0, nl_applied = 0;
        if (ok) {
            if (lead) vq = point_allele_vq(p, lc.qsum, cfg, nl_applied);
            else if (scan) sc.rmxn[item] = point_allele_rmxn(p, position, ref_allele, alt, cfg, ex) ? 1 : 0;
            else if (p.allele_support > 0) {
                const SbStats st = sb_stats_of(which, p.cov, p.sup, cfg.sb_noise, cfg.sb_model, cfg.sb_min_vf);
                sc.vg[item][which] = st.vg; sc.fp[item][which] = st.fp;
            }
        }
        __syncthreads();
        if (ok && lead && vq >= cfg.min_vq) {   // AlleleCaller.IsCallable
            SbResult sb;
            sb.bias = 0; sb.gatk = 0; sb.acceptable = false; sb.var_both = false; sb.cov_both = false;   // new BiasResults()
            if (p.allele_support > 0) {
                SbStats s3[3];
#pragma unroll
                for (int w = 0; w < 3; w++) {
                    int s_, c_;
                    sb_inputs_of(w, p.cov, p.sup, s_, c_);
                    s3[w].support = s_; s3[w].coverage = c_; s3[w].vg = sc.vg[item][w]; s3[w].fp = sc.fp[item][w]; s3[w].fn = 0;
                }
                sb = strand_bias_combine(s3[0], s3[1], s3[2], (double)cfg.sb_acceptance);
            }
            pb2_call_record rec;
            point_allele_finish(p, vq, nl_applied, sb, sc.rmxn[item] != 0, lc.qsum, position, ref_allele, alt, cfg, ex, rec);
            called = true;
            const unsigned long long slot = atomicAdd(out.var_count, 1ull);
            if ((int64_t)slot < out.var_capacity) store_record(out.var_records + slot, rec);
        }
    }
    if (have_item && lead && out.ref_records != nullptr) {
        const bool variant_called = ((cand_mask & 0x100) != 0) || called;
        const bool emit = cfg.output_gvcf && !variant_called && (cfg.have_intervals || (cand_mask & 0x200)) && (ex.chr_len == 0 || position <= ex.chr_len);
        if (emit) {
            pb2_call_record r;
            score_point_allele(lc, position, ref_allele, ref_allele, gapped, cfg, ex, r);
            store_record(out.ref_records + locus, r);
        }
        out.ref_valid[locus] = emit ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------ the hot kernel: vertical counters
// Point alleles (Reference, SNV) only ever read anchor-summed counts (CoverageCalculator.CalculateSinglePoint, RegionState.GetAllCandidates), so
// this kernel counts 18 rows (allele x direction) [+ 8 collapsed-read types] instead of 198 bins and keeps them in REGISTERS as bit-sliced
// ("vertical") counters: plane P[i] holds bit i of all row counters, one row per bit position. An entry is a one-hot word 1 << row; sixteen
// entries are added with a carry-save adder tree (two LOP3 per 3:2 compressor), the last carry ripples through the high planes. No shared
// memory, no atomics, no read-modify-write chain — the loop is pure ALU work next to two (three with collapsed reads) coalesced 16-byte
// loads per lane per step, and the anchor plane is not read at all. The 198-bin kernel above remains for pb2_get_counts (IAlleleSource).
constexpr int kVRows = 24;   // bit (direction * 8 + staged allele) for the 18 used rows; bits 24..31 = ReadCollapsedType 0..7

__device__ __forceinline__ void csa(uint32_t& sum, uint32_t& carry, uint32_t a, uint32_t b, uint32_t c) {
    carry = (a & b) | (a & c) | (b & c);
    sum = a ^ b ^ c;
}
// Rows of four entries after the quality rule, byte-parallel. Staged code bytes carry allele' | direction << 3 in their low 5 bits with N staged
// as 7, so row = direction * 8 + allele' is the low 5 bits as they are and `q < minBQ -> N` (RegionStateManager.cs:180-181) is an OR with 7.
// The ALU pipe (LOP3/SHF) is this kernel's bottleneck: the subtract and the x7 are written as multiplies to run on the FMA pipe.
__device__ __forceinline__ uint32_t rows_of_word(uint32_t c4, uint32_t q4, uint32_t neg_minbq4, uint32_t one) {
    // staged q carries bit 7: after the subtract bit 7 of a byte is set <=> q >= minBQ. Written as q * one + (-minBQ) with a run-time `one` so that it
    // issues on the FMA pipe (IMAD) instead of the saturated ALU pipe (IADD3).
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(q4), "r"(one), "r"(neg_minbq4));
    uint32_t ok;                                                    // PRMT with sign replication (selector nibbles 8..b; __byte_perm masks them off):
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(ok) : "r"(d), "r"(0u), "r"(0xba98u));   // 0xFF where q >= minBQ, 0x00 where the base becomes N
    return c4 | (~ok & 0x07070707u);                                // the flag bits 5-7 stay in the bytes: every consumer looks at the low 5 (or 3) bits only
}
template <bool kCollapsed>
__device__ __forceinline__ uint32_t onehot(uint32_t rows4, uint32_t a4, int k) {
    // byte k to the low bits with a high-multiply (FMA pipe) instead of a shift (ALU pipe); the variable shift uses the low 5 bits
    const uint32_t sh = k == 0 ? rows4 : __umulhi(rows4, 1u << (32 - 8 * k));
    uint32_t v = 1u << (sh & 31u);
    if (kCollapsed) {
        // CollapsedRegionState.AddCollapsedReadCount (:28-44): any entry not counted as N, typed reads only
        const uint32_t ct = (a4 >> (8 * k + 4)) & 0xfu;
        const bool usable = (sh & 7u) != kStagedN;
        if (ct != 0 && usable) {
            v |= 1u << (kVRows + ct - 1);
            if (ct - 1 == 4 || ct - 1 == 6) v |= 1u << (kVRows + 2);
            else if (ct - 1 == 5 || ct - 1 == 7) v |= 1u << (kVRows + 3);
        }
    }
    return v;
}
// counter of row r out of the planes: Σ_i bit r of P[i] << i  (mask on the ALU pipe, the shift-accumulate as a multiply on the FMA pipe)
template <int NP>
__device__ __forceinline__ int vcount_row(const uint32_t (&P)[NP], int r) {
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const uint32_t t = P[i] & (1u << r);
        v += (i >= r) ? t * (1u << (i - r)) : __umulhi(t, 1u << (32 - (r - i)));
    }
    return (int)v;
}

template <int NP, bool kWantQsum, bool kCollapsed, int kCtasPerSm, int kPf = 0>
__global__ void __launch_bounds__(256, kCtasPerSm)
pileup_vcount_score_kernel(const __grid_constant__ TilePileup in, const __grid_constant__ HotInputsExtra ex, const __grid_constant__ HotOutputs out,
                           const __grid_constant__ DeviceConfig cfg, int* __restrict__ tile_counter) {
    __shared__ int s_tile[8];
    __shared__ double q_lut[kWantQsum ? 256 : 1];
    __shared__ __align__(16) PendingLocus s_pend[kCtaPending];
    __shared__ int s_pend_n;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_pend_n = 0;
    if (kWantQsum) {
        for (int q = threadIdx.x; q < 256; q += blockDim.x) q_lut[q] = pow(10.0, (double)((float)(-q) / 10.0f));  // RegionStateManager.cs:191 (float exponent)
    }
    __syncthreads();
    const uint32_t neg_minbq4 = 0u - (uint32_t)cfg.min_bq * 0x01010101u;
    const uint32_t one = cfg.one;   // 1, read from the constant bank: opaque to the compiler, costs no register

    while (true) {
        if (lane == 0) s_tile[warp] = atomicAdd(tile_counter, 1);
        __syncwarp();
        const int tile = s_tile[warp];
        __syncwarp();
        if (tile >= in.n_tiles) break;

        const int64_t locus = (int64_t)tile * kTileLoci + lane;
        const bool have_locus = locus < in.n_loci;
        const int depth = have_locus ? in.depth[locus] : 0;
        const int nchunks = (depth + kChunk - 1) / kChunk;
        const int max_chunks = __reduce_max_sync(0xffffffffu, nchunks);
        int64_t base = in.tile_base[tile];
        double qsum = 0.0;
        int extra_pad = 0;   // whole PAD chunks counted by lanes that ran out of entries before the longest lane of the tile
        uint32_t P[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) P[i] = 0;

        // One 16-entry chunk into the planes below weight 16; returns the weight-16 carry (rippled by the caller, every second chunk in phase 1).
        auto add_chunk = [&](const uint4& wc, const uint4& wq, const uint4& wa) -> uint32_t {
            const uint32_t r0 = rows_of_word(wc.x, wq.x, neg_minbq4, one), r1 = rows_of_word(wc.y, wq.y, neg_minbq4, one),
                           r2 = rows_of_word(wc.z, wq.z, neg_minbq4, one), r3 = rows_of_word(wc.w, wq.w, neg_minbq4, one);
            // weight 1: P[0] + 16 one-hot words -> P[0] and eight weight-2 carries
            uint32_t s0, s1, s2, s3, s4, t0, t1, k0, k1, k2, k3, k4, k5, k6, k7;
            csa(s0, k0, onehot<kCollapsed>(r0, wa.x, 0), onehot<kCollapsed>(r0, wa.x, 1), onehot<kCollapsed>(r0, wa.x, 2));
            csa(s1, k1, onehot<kCollapsed>(r0, wa.x, 3), onehot<kCollapsed>(r1, wa.y, 0), onehot<kCollapsed>(r1, wa.y, 1));
            csa(s2, k2, onehot<kCollapsed>(r1, wa.y, 2), onehot<kCollapsed>(r1, wa.y, 3), onehot<kCollapsed>(r2, wa.z, 0));
            csa(s3, k3, onehot<kCollapsed>(r2, wa.z, 1), onehot<kCollapsed>(r2, wa.z, 2), onehot<kCollapsed>(r2, wa.z, 3));
            csa(s4, k4, onehot<kCollapsed>(r3, wa.w, 0), onehot<kCollapsed>(r3, wa.w, 1), onehot<kCollapsed>(r3, wa.w, 2));
            csa(t0, k5, s0, s1, s2);
            csa(t1, k6, s3, s4, onehot<kCollapsed>(r3, wa.w, 3));
            csa(P[0], k7, P[0], t0, t1);
            // weight 2: P[1] + 8 carries -> P[1] and four weight-4 carries
            uint32_t u0, u1, u2, m0, m1, m2, m3;
            csa(u0, m0, k0, k1, k2);
            csa(u1, m1, k3, k4, k5);
            csa(u2, m2, k6, k7, P[1]);
            csa(P[1], m3, u0, u1, u2);
            // weight 4: P[2] + 4 carries -> P[2] and two weight-8 carries
            uint32_t w0, n0, n1;
            csa(w0, n0, m0, m1, m2);
            csa(P[2], n1, w0, m3, P[2]);
            // weight 8: P[3] + 2 carries -> P[3] and one weight-16 carry
            uint32_t carry;
            csa(P[3], carry, n0, n1, P[3]);
            if (kWantQsum) {   // Σ 10^(-q/10f) over entries counted as A/C/G/T (Deletion entries carry no base quality, :191)
                const uint32_t cw[4] = {wc.x, wc.y, wc.z, wc.w}, qw[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
                for (int k = 0; k < kChunk; k++) {
                    const uint32_t code = (cw[k >> 2] >> ((k & 3) * 8)) & 0xffu, q = (qw[k >> 2] >> ((k & 3) * 8)) & kQualMask;
                    if ((code & 7) < AT_N && (int)q >= cfg.min_bq) qsum += q_lut[q];
                }
            }
            return carry;
        };
        auto ripple = [&](uint32_t carry, int from) {
#pragma unroll
            for (int i = 4; i < NP; i++) if (i >= from) { const uint32_t t = P[i] & carry; P[i] ^= carry; carry = t; }
        };

        // ---- phase 1: the steps every lane of the tile takes part in (uniform: slot == lane, 512 contiguous bytes per plane per step, no ballots,
        // immediate address offsets). Chunk c+1 is in flight in registers and chunk c+2 is prefetched into L2 while chunk c is counted. Two chunks
        // per iteration so that their weight-16 carries meet in one more compressor and the ripple through the high planes runs half as often.
        const int min_chunks = __reduce_min_sync(0xffffffffu, nchunks);
        {
            constexpr int kStep = kTileLoci * kChunk;   // bytes of one full step in one plane; code and quality share a plane: 2 * kStep per step there
            const uint8_t* pc = in.cq + 2 * base + lane * kChunk;
            const uint8_t* pa = in.anch + base + lane * kChunk;
            uint4 nc = make_uint4(0, 0, 0, 0), nq = nc, na = nc;
            if (min_chunks > 0) { nc = ldg_stream(pc); nq = ldg_stream(pc + kStep); if (kCollapsed) na = ldg_stream(pa); }
            // chunk j of the current iteration: the next chunk goes into the registers, the one after it into L2, then this one is counted
            auto step = [&](int j, bool more) -> uint32_t {
                const uint4 wc = nc, wq = nq, wa = na;
                if (more) { nc = ldg_stream(pc + (j + 1) * 2 * kStep); nq = ldg_stream(pc + (j + 1) * 2 * kStep + kStep); if (kCollapsed) na = ldg_stream(pa + (j + 1) * kStep); }
                // kPf steps ahead into L2 (the planes carry slack for the prefetches past the last tile)
                if (kPf > 0) { prefetch_l2(pc + (j + kPf) * 2 * kStep); prefetch_l2(pc + (j + kPf) * 2 * kStep + kStep); if (kCollapsed) prefetch_l2(pa + (j + kPf) * kStep); }
                return add_chunk(wc, wq, wa);
            };
            auto advance = [&](int n) { pc += n * 2 * kStep; if (kCollapsed) pa += n * kStep; };
            int c = 0;
            // four chunks per iteration: the weight-16 carries pair up into weight 32, those into weight 64, and only that one ripples through the high
            // planes (a quarter of the ripples); the plane pointers advance once, the loads use immediate offsets
            for (; c + 4 <= min_chunks; c += 4) {
                uint32_t ca = step(0, true);
                uint32_t cb = step(1, true);
                uint32_t c32a, c32b, c64;
                csa(P[4], c32a, ca, cb, P[4]);
                ca = step(2, true);
                cb = step(3, c + 4 < min_chunks);
                csa(P[4], c32b, ca, cb, P[4]);
                csa(P[5], c64, c32a, c32b, P[5]);
                ripple(c64, 6);
                advance(4);
            }
            for (; c + 2 <= min_chunks; c += 2) {
                const uint32_t ca = step(0, true);
                const uint32_t cb = step(1, c + 2 < min_chunks);
                uint32_t c32;
                csa(P[4], c32, ca, cb, P[4]);
                ripple(c32, 5);
                advance(2);
            }
            if (c < min_chunks) ripple(step(0, false), 4);
            base += (int64_t)min_chunks * kStep;
        }

        // ---- phase 2: the ragged end of the tile. Lanes that ran out of entries count PAD chunks (subtracted at read-out).
        if (min_chunks < max_chunks) {
            const uint4 pad_c = make_uint4(kPadCode4, kPadCode4, kPadCode4, kPadCode4), pad_q = make_uint4(~0u, ~0u, ~0u, ~0u), pad_a = make_uint4(0, 0, 0, 0);
            uint4 nc = pad_c, nq = pad_q, na = pad_a;
            {
                const bool active = min_chunks < nchunks;
                const unsigned m = __ballot_sync(0xffffffffu, active);
                if (active) {
                    const int64_t o = base + (int64_t)__popc(m & ((1u << lane) - 1)) * kChunk;
                    nc = ldg_stream(in.cq + base + o); nq = ldg_stream(in.cq + base + o + (int64_t)__popc(m) * kChunk);
                    if (kCollapsed) na = ldg_stream(in.anch + o);
                }
                base += (int64_t)__popc(m) * kChunk;
            }
#pragma unroll 1
            for (int c = min_chunks; c < max_chunks; c++) {
                const uint4 wc = nc, wq = nq, wa = na;
                if (c >= nchunks) extra_pad += kChunk;
                nc = pad_c; nq = pad_q; na = pad_a;
                {
                    const bool active = (c + 1) < nchunks;
                    const unsigned m = __ballot_sync(0xffffffffu, active);
                    if (active) {
                        const int64_t o = base + (int64_t)__popc(m & ((1u << lane) - 1)) * kChunk;
                        nc = ldg_stream(in.cq + base + o); nq = ldg_stream(in.cq + base + o + (int64_t)__popc(m) * kChunk);
                        if (kCollapsed) na = ldg_stream(in.anch + o);
                    }
                    base += (int64_t)__popc(m) * kChunk;
                }
                ripple(add_chunk(wc, wq, wa), 4);
            }
        }

        // ---- read the vertical counters out
        int cnt[kNumAlleles][kNumDirs];
        int any = 0;
        const int npad = (have_locus ? in.pad[locus] : 0) + extra_pad;
#pragma unroll
        for (int a = 0; a < kNumAlleles; a++)
#pragma unroll
            for (int d = 0; d < kNumDirs; d++) {
                int v = vcount_row<NP>(P, d * 8 + (a == AT_N ? (int)kStagedN : a));
                if (a == AT_N && d == DIR_F) v -= npad;
                cnt[a][d] = v;
                any += v;
            }
        if (kCollapsed) {
            if (out.collapsed_out != nullptr && have_locus) {
#pragma unroll
                for (int t = 0; t < kNumCollapsed; t++) out.collapsed_out[locus * kNumCollapsed + t] = vcount_row<NP>(P, kVRows + t);
            }
        }
        if (!have_locus) continue;
        const int ref_allele = allele_of_base(in.ref_base[locus]);
        finish_locus(cnt, qsum, any, locus, ref_allele, in, ex, out, cfg, s_pend, &s_pend_n);
    }

    // ---- the CTA's queued loci: 4 lanes per locus, all 256 threads at once. Other CTAs of this SM are still counting, so this FP64 latency
    // chain overlaps their work; only the last CTAs' pass is exposed at the end of the kernel.
    __syncthreads();
    {
        const int n = min(s_pend_n, kCtaPending);
        const int item = threadIdx.x >> 2;
        score_queued_locus(item < n ? &s_pend[item] : nullptr, threadIdx.x & 3, in, ex, out, cfg);
    }
}

// One thread per queued locus: the full ProcessVariant + genotype chain for its SNV candidates in (ref, alt) order (AlleleCaller.cs:172-176),
// then the reference allele if nothing was called there (:146-147).
__global__ void __launch_bounds__(128) score_pending_kernel(const __grid_constant__ TilePileup in, const __grid_constant__ HotInputsExtra ex,
                                                            const __grid_constant__ HotOutputs out, const __grid_constant__ DeviceConfig cfg) {
    const unsigned long long n = min(*out.pending_count, (unsigned long long)out.pending_capacity);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const PendingLocus pl = out.pending[i];
        const int64_t locus = pl.locus;
        LocusCounts lc;
#pragma unroll
        for (int a = 0; a < kNumAlleles; a++)
#pragma unroll
            for (int d = 0; d < kNumDirs; d++) lc.c[a][d] = pl.c[a * kNumDirs + d];
        lc.qsum = pl.qsum;
        const int ref_allele = allele_of_base(in.ref_base[locus]);
        const int position = in.positions ? in.positions[locus] : in.first_position + (int)locus;
        bool variant_called = (pl.cand_mask & 0x100) != 0;
        const int order[4] = {AT_A, AT_C, AT_G, AT_T};
#pragma unroll 1
        for (int oi = 0; oi < 4; oi++) {
            const int alt = order[oi];
            if (!((pl.cand_mask >> alt) & 1)) continue;
            pb2_call_record r;
            if (score_point_allele(lc, position, ref_allele, alt, pl.gapped, cfg, ex, r)) {
                variant_called = true;
                const unsigned long long slot = atomicAdd(out.var_count, 1ull);
                if ((int64_t)slot < out.var_capacity) store_record(out.var_records + slot, r);
            }
        }
        if (out.ref_records != nullptr) {
            const bool emit = cfg.output_gvcf && !variant_called && (cfg.have_intervals || (pl.cand_mask & 0x200)) && (ex.chr_len == 0 || position <= ex.chr_len);
            if (emit) {
                pb2_call_record r;
                score_point_allele(lc, position, ref_allele, ref_allele, pl.gapped, cfg, ex, r);
                store_record(out.ref_records + locus, r);
            }
            out.ref_valid[locus] = emit ? 1 : 0;
        }
    }
}

size_t hot_kernel_smem_bytes(bool narrow, bool collapsed) {
    if (narrow) return (size_t)(kNumBins + kNumAlleles * kNumDirs) * kNarrowThreads + 256 * sizeof(double);
    const int rows = kNumBins + (collapsed ? kNumCollapsed : 0);
    return (size_t)rows * kHotThreads * sizeof(uint16_t) + 256 * sizeof(double);
}

// ================================================================================================ PNIB16: direction-split, nibble-packed pileup
// The vertical-counter kernel above spends its time on the integer pipe: a shift per entry to make the one-hot row word and 15 compressors per
// 16 entries, because allele AND direction select the row. Here the direction is a property of the lane (sub-locus = (locus, direction)) and the
// allele a 4-bit code, so ONE byte-permute turns four entries into four one-hot bytes (PRMT as a table lookup: selector nibble = allele code,
// table = {1, 2, 4, 8, 0, 16, 0, 0}), the quality rule is an AND with the byte mask of `q >= minBQ`, and four words per chunk go through the
// compressors instead of sixteen. Counters are vertical over (byte lane, allele): 4 x 8 = 32 of them per plane. 1.5 bytes per entry instead of 2.
__device__ __forceinline__ uint32_t spread16(uint32_t x) {   // bit i -> bit 2i
    x &= 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
__device__ __forceinline__ int nib_step_bytes(int active) { return (24 * active + 15) & ~15; }
constexpr uint32_t kNibPadCode = 6;   // table entry 0: contributes nothing

// one warp per PTILE32 tile, lane = locus: calls f(code chunk, quality chunk) for the lane's own chunks, in order (the addressing of tile_scatter_kernel)
template <class F>
__device__ __forceinline__ void walk_ptile(const TilePileup& in, int64_t tile, int lane, F&& f) {
    const int64_t locus = tile * kTileLoci + lane;
    const int depth = locus < in.n_loci ? in.depth[locus] : 0;
    const int nchunks = (depth + kChunk - 1) / kChunk;
    const int max_chunks = __reduce_max_sync(0xffffffffu, nchunks);
    int64_t base = in.tile_base[tile];
    for (int c = 0; c < max_chunks; c++) {
        const bool active = c < nchunks;
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (active) {
            const int64_t o = 2 * base + (int64_t)__popc(m & ((1u << lane) - 1)) * kChunk;
            const uint4 wc = *reinterpret_cast<const uint4*>(in.cq + o);
            const uint4 wq = *reinterpret_cast<const uint4*>(in.cq + o + (int64_t)__popc(m) * kChunk);
            f(wc, wq);
        }
        base += (int64_t)__popc(m) * kChunk;
    }
}

__global__ void nib_count_kernel(const __grid_constant__ TilePileup in, int32_t* __restrict__ nib_store, int32_t* __restrict__ nib_depth,
                                 int64_t* __restrict__ nib_tile_bytes, int32_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tile >= in.n_tiles) return;
    int st[2] = {0, 0}, dp[2] = {0, 0};
    bool stitched = false;
    walk_ptile(in, tile, lane, [&](const uint4& wc, const uint4& wq) {
        const uint32_t cw[4] = {wc.x, wc.y, wc.z, wc.w}, qw[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {   // four entries at a time; every mask below lives in bit 3 of its byte
            const uint32_t c4 = cw[k], y = ~qw[k];
            const uint32_t n7 = ((c4 & 0x07070707u) + 0x01010101u) & 0x08080808u;                                  // staged allele code 7 (N or PAD)
            const uint32_t qff = (~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y) & 0x80808080u) >> 4;                   // quality byte 0xff
            const uint32_t pad = n7 & qff;
            const uint32_t dir_r = c4 & 0x08080808u, dir_s = (c4 & 0x10101010u) >> 1;
            const uint32_t valid = 0x08080808u & ~pad & ~dir_s;
            stitched |= (dir_s & ~pad) != 0;
            dp[1] += __popc(valid & dir_r); dp[0] += __popc(valid & ~dir_r);
            const uint32_t stored = valid & ~n7;
            st[1] += __popc(stored & dir_r); st[0] += __popc(stored & ~dir_r);
        }
    });
    const int64_t locus = tile * kTileLoci + lane;
    if (locus < in.n_loci) {
        nib_store[2 * locus] = st[0]; nib_store[2 * locus + 1] = st[1];
        nib_depth[2 * locus] = dp[0]; nib_depth[2 * locus + 1] = dp[1];
    }
    const int nc0 = (st[0] + kChunk - 1) / kChunk, nc1 = (st[1] + kChunk - 1) / kChunk;
    const int maxc = __reduce_max_sync(0xffffffffu, max(nc0, nc1));
    int64_t bytes_lo = 0, bytes_hi = 0;
    for (int c = 0; c < maxc; c++) {
        const unsigned b0 = __ballot_sync(0xffffffffu, nc0 > c), b1 = __ballot_sync(0xffffffffu, nc1 > c);
        bytes_lo += nib_step_bytes(__popc(b0 & 0xffffu) + __popc(b1 & 0xffffu));
        bytes_hi += nib_step_bytes(__popc(b0 >> 16) + __popc(b1 >> 16));
    }
    if (lane == 0) {
        nib_tile_bytes[2 * tile] = bytes_lo;
        if (2 * tile + 1 < in.n_nib_tiles) nib_tile_bytes[2 * tile + 1] = bytes_hi;
    }
    const int mx = __reduce_max_sync(0xffffffffu, max(st[0], st[1]));
    if (lane == 0) atomicMax(flags + 1, mx);
    if (__any_sync(0xffffffffu, stitched) && lane == 0) atomicOr(flags, 1);
}

constexpr int kNibScatterWarps = 4;
__global__ void __launch_bounds__(32 * kNibScatterWarps) nib_scatter_kernel(const __grid_constant__ TilePileup in, const int32_t* __restrict__ nib_store,
                                                                           const int64_t* __restrict__ nib_tile_base, uint8_t* __restrict__ nib) {
    __shared__ int32_t s_off[kNibScatterWarps][2][kNibMaxChunks];
    __shared__ uint32_t s_mask[kNibScatterWarps][2][kNibMaxChunks];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * kNibScatterWarps + warp;
    if (tile >= in.n_tiles) return;
    const int64_t locus = tile * kTileLoci + lane;
    int nst[2] = {0, 0};
    if (locus < in.n_loci) { nst[0] = nib_store[2 * locus]; nst[1] = nib_store[2 * locus + 1]; }
    const int nc0 = (nst[0] + kChunk - 1) / kChunk, nc1 = (nst[1] + kChunk - 1) / kChunk;
    const int maxc = min(__reduce_max_sync(0xffffffffu, max(nc0, nc1)), kNibMaxChunks);
    {   // per step: which of the sub-tile's 32 sub-loci are active (bit 2 * (lane & 15) + direction) and where the step starts
        int off_lo = 0, off_hi = 0;
        for (int c = 0; c < maxc; c++) {
            const unsigned b0 = __ballot_sync(0xffffffffu, nc0 > c), b1 = __ballot_sync(0xffffffffu, nc1 > c);
            const uint32_t m_lo = spread16(b0) | (spread16(b1) << 1), m_hi = spread16(b0 >> 16) | (spread16(b1 >> 16) << 1);
            if (lane == 0) { s_off[warp][0][c] = off_lo; s_off[warp][1][c] = off_hi; s_mask[warp][0][c] = m_lo; s_mask[warp][1][c] = m_hi; }
            off_lo += nib_step_bytes(__popc(m_lo));
            off_hi += nib_step_bytes(__popc(m_hi));
        }
    }
    __syncwarp();
    const int sub = lane >> 4;
    const int64_t tb = (2 * tile + sub < in.n_nib_tiles) ? nib_tile_base[2 * tile + sub] : 0;
    // per (lane, direction) a 32-entry ring of pending entries in shared memory, one byte per allele code and one per quality, laid out
    // [direction][word of 4 entries][lane] so that every access is bank-conflict free. Appending is two byte stores at a computed address (entries that
    // are not stored go to a per-lane dump word: no branches in the per-entry code); 16 entries leave as one chunk.
    __shared__ uint32_t s_al[kNibScatterWarps][2][8][32], s_q[kNibScatterWarps][2][8][32], s_dump[kNibScatterWarps][32];
    uint8_t* const al_bytes = reinterpret_cast<uint8_t*>(&s_al[warp][0][0][0]);
    uint8_t* const q_bytes = reinterpret_cast<uint8_t*>(&s_q[warp][0][0][0]);
    uint8_t* const dump = reinterpret_cast<uint8_t*>(&s_dump[warp][lane]);
    int fill[2] = {0, 0}, head[2] = {0, 0}, cidx[2] = {0, 0};
    auto flush = [&](int d, int n) {   // the oldest min(n, 16) pending entries of direction d become chunk cidx[d] of the sub-locus
        const int c = cidx[d];
        uint32_t a4[4], q4[4];
#pragma unroll
        for (int w = 0; w < 4; w++) { a4[w] = s_al[warp][d][(head[d] >> 2) + w][lane]; q4[w] = s_q[warp][d][(head[d] >> 2) + w][lane]; }
        if (n < kChunk) {   // the last, partial chunk: PAD the tail
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const int keep = min(max(n - 4 * w, 0), 4);
                const uint32_t km = keep >= 4 ? 0xffffffffu : ((1u << (8 * keep)) - 1u);
                a4[w] = (a4[w] & km) | (0x06060606u & ~km);
                q4[w] = q4[w] | ~km;
            }
        }
        uint32_t nib16[4];
#pragma unroll
        for (int w = 0; w < 4; w++) nib16[w] = __byte_perm(a4[w] | (a4[w] >> 4), 0, 0x4420) & 0xffffu;   // bytes b0|b1<<4, b2|b3<<4
        if (c < kNibMaxChunks) {
            const uint32_t m = s_mask[warp][sub][c];
            const int j = 2 * (lane & 15) + d;
            const int active = __popc(m), rank = __popc(m & ((1u << j) - 1));
            uint8_t* dst = nib + tb + s_off[warp][sub][c];
            *reinterpret_cast<uint4*>(dst + 16 * rank) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
            *reinterpret_cast<uint2*>(dst + 16 * active + 8 * rank) = make_uint2(nib16[0] | (nib16[1] << 16), nib16[2] | (nib16[3] << 16));
        }
        head[d] = (head[d] + kChunk) & 31;
        fill[d] -= min(n, kChunk);
        cidx[d] = c + 1;
    };
    walk_ptile(in, tile, lane, [&](const uint4& wc, const uint4& wq) {
        const uint32_t cw[4] = {wc.x, wc.y, wc.z, wc.w}, qw[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
        for (int k = 0; k < kChunk; k++) {
            const uint32_t c = (cw[k >> 2] >> ((k & 3) * 8)) & 0xffu, q = (qw[k >> 2] >> ((k & 3) * 8)) & 0xffu;
            const uint32_t al = c & 7u, dir = (c >> 3) & 1u;
            const bool take = (al != kStagedN) & (((c >> 4) & 1u) == 0);   // PADs and N bases are not stored (N = counted - stored alleles)
            const int f = (dir ? head[1] + fill[1] : head[0] + fill[0]) & 31;
            const int off = (((int)dir * 8 + (f >> 2)) * 32 + lane) * 4 + (f & 3);
            *(take ? al_bytes + off : dump) = (uint8_t)al;
            *(take ? q_bytes + off : dump + 1) = (uint8_t)q;
            fill[0] += (take & (dir == 0)) ? 1 : 0;
            fill[1] += (take & (dir == 1)) ? 1 : 0;
        }
        if (fill[0] >= kChunk) flush(0, kChunk);
        if (fill[1] >= kChunk) flush(1, kChunk);
    });
    if (fill[0] > 0) flush(0, fill[0]);
    if (fill[1] > 0) flush(1, fill[1]);
}

// counter of (byte lane b, allele bit a) summed over the four byte lanes
template <int NP>
__device__ __forceinline__ int nib_count_allele(const uint32_t (&P)[NP], int a) {
    int v = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) v += __popc(P[i] & (0x01010101u << a)) << i;   // plane i: how many of the four byte lanes have bit i of their counter set
    return v;
}

template <int NP, bool kPair>
__global__ void __launch_bounds__(256, 4)
pileup_nib_score_kernel(const __grid_constant__ TilePileup in, const __grid_constant__ HotInputsExtra ex, const __grid_constant__ HotOutputs out,
                        const __grid_constant__ DeviceConfig cfg, int* __restrict__ tile_counter) {
    __shared__ __align__(16) PendingLocus s_pend[kCtaPending];
    __shared__ int s_pend_n;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_pend_n = 0;
    __syncthreads();
    const uint32_t neg_minbq4 = 0u - (uint32_t)cfg.min_bq * 0x01010101u;
    const uint32_t one = cfg.one;
    const uint32_t t_lo = 0x08040201u, t_hi = 0x00001000u;   // allele code -> one-hot byte: A 1, G 2, C 4, T 8, (4) 0, Deletion 16, PAD 0, (7) 0
    const unsigned lt = (1u << lane) - 1;

    auto grab = [&]() -> int {
        int t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    int next_pair = grab();
    while (true) {
        // a warp takes two sub-tiles (32 loci) at a time: each is counted with one (locus, direction) per lane, then the 2 x 16 loci are scored with one
        // locus per lane, so that the per-locus work (candidate screening; in gVCF mode the reference allele's q-score / genotype chain) runs on full warps
        const int tile_pair = next_pair;
        if ((kPair ? 2 : 1) * tile_pair >= in.n_nib_tiles) break;
        uint32_t keep[kNumAlleles];   // this lane's locus: counts packed as forward | reverse << 16
#pragma unroll
        for (int a = 0; a < kNumAlleles; a++) keep[a] = 0;
#pragma unroll 1
        for (int half = 0; half < (kPair ? 2 : 1); half++) {
        const int tile = (kPair ? 2 : 1) * tile_pair + half;
        if (tile >= in.n_nib_tiles) break;
        const int64_t locus = (int64_t)tile * kNibLoci + (lane >> 1);
        const bool have_locus = locus < in.n_loci;
        const int64_t sub = 2 * locus + (lane & 1);
        const int store = have_locus ? in.nib_store[sub] : 0;
        const int nchunks = (store + kChunk - 1) / kChunk;
        const int max_chunks = __reduce_max_sync(0xffffffffu, nchunks);
        const uint8_t* p = in.nib + in.nib_tile_base[tile];
        uint32_t P[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) P[i] = 0;

        // four one-hot byte words of a chunk (quality rule applied) into the planes below weight 4; returns the weight-4 carry
        auto add_chunk = [&](const uint2& wc, const uint4& wq) -> uint32_t {
            const uint32_t sel[4] = {wc.x, wc.x >> 16, wc.y, wc.y >> 16};
            const uint32_t qv[4] = {wq.x, wq.y, wq.z, wq.w};
            uint32_t x[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t oh, d, ok;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(oh) : "r"(t_lo), "r"(t_hi), "r"(sel[k]));
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(qv[k]), "r"(one), "r"(neg_minbq4));
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(ok) : "r"(d), "r"(0u), "r"(0xba98u));   // 0xFF where q >= minBQ (RegionStateManager.cs:180-181)
                x[k] = oh & ok;
            }
            uint32_t s0, k0, k1, m0;
            csa(s0, k0, x[0], x[1], x[2]);
            csa(P[0], k1, P[0], s0, x[3]);
            csa(P[1], m0, k0, k1, P[1]);
            return m0;
        };
        auto ripple = [&](uint32_t carry, int from) {
#pragma unroll
            for (int i = 2; i < NP; i++) if (i >= from) { const uint32_t t = P[i] & carry; P[i] ^= carry; carry = t; }
        };

        // two chunks are counted per iteration while the next two are already on their way into registers (the kernel is bound by the bytes it keeps in
        // flight, not by the integer pipe: 4 chunks x 24 B per lane outstanding)
        auto fetch = [&](int c, uint2& dc, uint4& dq) {
            const bool active = c < nchunks;
            const unsigned m = __ballot_sync(0xffffffffu, active);
            const int n_act = __popc(m);
            dc = make_uint2(kNibPadCode * 0x11111111u, kNibPadCode * 0x11111111u);
            dq = make_uint4(~0u, ~0u, ~0u, ~0u);
            if (active) {
                const int rank = __popc(m & lt);
                dq = ldg_stream(p + 16 * rank);
                const uint8_t* pcode = p + 16 * n_act + 8 * rank;
                asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.u32 {%0,%1}, [%2];" : "=r"(dc.x), "=r"(dc.y) : "l"(pcode));
            }
            p += nib_step_bytes(n_act);
        };
        uint2 c0, c1, c2, c3;
        uint4 q0, q1, q2, q3;
        int c = 0;
        // ---- phase 1: the steps every sub-locus of the sub-tile takes part in: slot == lane, 768 contiguous bytes per step (512 of quality chunks, 256
        // of code chunks), no ballots, immediate offsets; four chunks per iteration, the next pair always in flight (ping-pong, no register moves)
        const int uniform = __reduce_min_sync(0xffffffffu, nchunks) & ~3;
        if (uniform > 0) {
            constexpr int kFull = 768;
            const uint8_t* pq = p + 16 * lane;
            const uint8_t* pk = p + 512 + 8 * lane;
            auto ld = [&](int j, uint2& dc, uint4& dq) {
                dq = ldg_stream(pq + j * kFull);
                asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.u32 {%0,%1}, [%2];" : "=r"(dc.x), "=r"(dc.y) : "l"(pk + j * kFull));
            };
            ld(0, c0, q0);
            ld(1, c1, q1);
            for (; c < uniform; c += 4) {
                ld(2, c2, q2);
                ld(3, c3, q3);
                uint32_t ca = add_chunk(c0, q0);
                uint32_t cb = add_chunk(c1, q1);
                uint32_t c8a, c8b, c16;
                csa(P[2], c8a, ca, cb, P[2]);
                if (c + 4 < uniform) { ld(4, c0, q0); ld(5, c1, q1); }
                ca = add_chunk(c2, q2);
                cb = add_chunk(c3, q3);
                csa(P[2], c8b, ca, cb, P[2]);
                csa(P[3], c16, c8a, c8b, P[3]);
                ripple(c16, 4);
                pq += 4 * kFull; pk += 4 * kFull;
            }
            p += (int64_t)uniform * kFull;
        }
        // ---- phase 2: the ragged end (sub-loci that ran out come back as PAD chunks without touching memory; max_chunks is uniform across the warp)
        if (c < max_chunks) {
            fetch(c, c0, q0);
            fetch(c + 1, c1, q1);
            for (; c < max_chunks; c += 2) {
                fetch(c + 2, c2, q2);
                fetch(c + 3, c3, q3);
                const uint32_t ca = add_chunk(c0, q0);
                const uint32_t cb = add_chunk(c1, q1);
                uint32_t c8;
                csa(P[2], c8, ca, cb, P[2]);
                ripple(c8, 3);
                c0 = c2; q0 = q2; c1 = c3; q1 = q3;
            }
        }

        // ---- counts of this (locus, direction); the partner lane holds the other direction
        int mine[kNumAlleles];
        mine[AT_A] = nib_count_allele<NP>(P, 0);
        mine[AT_G] = nib_count_allele<NP>(P, 1);
        mine[AT_C] = nib_count_allele<NP>(P, 2);
        mine[AT_T] = nib_count_allele<NP>(P, 3);
        mine[AT_DEL] = nib_count_allele<NP>(P, 4);
        const int counted = have_locus ? in.nib_depth[sub] : 0;
        mine[AT_N] = counted - (mine[AT_A] + mine[AT_G] + mine[AT_C] + mine[AT_T] + mine[AT_DEL]);   // N bases and bases below the quality bar
#pragma unroll
        for (int a = 0; a < kNumAlleles; a++) {
            const uint32_t other = (uint32_t)__shfl_xor_sync(0xffffffffu, mine[a], 1);
            const uint32_t packed = (uint32_t)mine[a] | (other << 16);                       // valid in the even (forward) lanes
            if (kPair) {
                const uint32_t got = __shfl_sync(0xffffffffu, packed, 2 * (lane & 15));      // locus (lane & 15) of this sub-tile
                if ((lane >> 4) == half) keep[a] = got;
            } else keep[a] = packed;
        }
        }   // half

        // the next piece of work is taken before this one is finished; in gVCF mode, where finishing means a q-score / strand-bias / genotype chain per
        // locus during which the warp has no loads in flight, the head of the next pair (2 x 8 KB) is pulled into L2 meanwhile
        next_pair = grab();
        if (kPair && cfg.output_gvcf && cfg.tune_prefetch != 8) {   // VCF mode: the prefetch costs 0.186 against 0.177 ms (measured)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int t = 2 * next_pair + h;
                if (t < in.n_nib_tiles) {
                    const int64_t b0 = in.nib_tile_base[t], b1 = in.nib_tile_base[t + 1];   // [n_nib_tiles + 1] entries
                    if (b0 + 128 * lane < b1) prefetch_l2(in.nib + b0 + 128 * lane);
                    if (b0 + 128 * (32 + lane) < b1) prefetch_l2(in.nib + b0 + 128 * (32 + lane));
                }
            }
        }
        // kPair (gVCF: every locus scores its reference allele): one locus per lane over both sub-tiles; otherwise the even lanes finish their own locus
        const int64_t locus = kPair ? ((int64_t)2 * tile_pair + (lane >> 4)) * kNibLoci + (lane & 15) : (int64_t)tile_pair * kNibLoci + (lane >> 1);
        if (kPair ? (2 * tile_pair + (lane >> 4) >= in.n_nib_tiles || locus >= in.n_loci) : ((lane & 1) || locus >= in.n_loci)) continue;
        int cnt[kNumAlleles][kNumDirs];
        int any = 0;
#pragma unroll
        for (int a = 0; a < kNumAlleles; a++) {
            cnt[a][DIR_F] = (int)(keep[a] & 0xffffu); cnt[a][DIR_R] = (int)(keep[a] >> 16); cnt[a][DIR_S] = 0;
            any += cnt[a][DIR_F] + cnt[a][DIR_R];
        }
        const int ref_allele = allele_of_base(in.ref_base[locus]);
        finish_locus(cnt, 0.0, any, locus, ref_allele, in, ex, out, cfg, s_pend, &s_pend_n);
    }

    __syncthreads();
    {
        const int n = min(s_pend_n, kCtaPending);
        const int item = threadIdx.x >> 2;
        score_queued_locus(item < n ? &s_pend[item] : nullptr, threadIdx.x & 3, in, ex, out, cfg);
    }
}

// table[cov][a] = Poisson.Cdf(a - 1, targetLOD * cov) with the float product of SomaticGenotypeQualityCalculator.cs:33 (see somatic_gq); behind it
// int32 gq[cov][a] = the GQ of a homozygous genotype whose variant q-score is capped_vq (the same statements as somatic_gq's tail)
__global__ void gq_tail_fill_kernel(double* __restrict__ table, float target_lod, double p1_capped, int min_gq, int max_gq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kGqTailMaxCov * kGqTailMaxA) return;
    const int cov = i / kGqTailMaxA, a = i % kGqTailMaxA;
    const float expected = target_lod * (float)cov;
    const double p2 = a >= 1 ? pisces_poisson_cdf((double)(a - 1), (double)expected) : 0.0;
    table[i] = p2;
    const double raw = -10 * log10(p1_capped + p2);
    double q = fmin((double)max_gq, raw);
    q = fmax(q, (double)min_gq);
    reinterpret_cast<int*>(table + kGqTailMaxCov * kGqTailMaxA)[i] = (int)rint(q);
}
cudaError_t launch_gq_tail_fill(double* table, float target_lod, double p1_capped, int min_gq, int max_gq, cudaStream_t stream) {
    gq_tail_fill_kernel<<<(kGqTailMaxCov * kGqTailMaxA + 255) / 256, 256, 0, stream>>>(table, target_lod, p1_capped, min_gq, max_gq);
    return cudaGetLastError();
}

cudaError_t launch_nib_count(const TilePileup& in, int32_t* nib_store, int32_t* nib_depth, int64_t* nib_tile_bytes, int32_t* flags, cudaStream_t stream) {
    if (in.n_tiles == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)(((int64_t)in.n_tiles * 32 + 255) / 256);
    nib_count_kernel<<<blocks, 256, 0, stream>>>(in, nib_store, nib_depth, nib_tile_bytes, flags);
    return cudaGetLastError();
}
cudaError_t launch_nib_scatter(const TilePileup& in, const int32_t* nib_store, const int64_t* nib_tile_base, uint8_t* nib, cudaStream_t stream) {
    if (in.n_tiles == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((in.n_tiles + kNibScatterWarps - 1) / kNibScatterWarps);
    nib_scatter_kernel<<<blocks, 32 * kNibScatterWarps, 0, stream>>>(in, nib_store, nib_tile_base, nib);
    return cudaGetLastError();
}

cudaError_t launch_hot_kernel(const TilePileup& in, const HotInputsExtra& ex, const HotOutputs& out, const DeviceConfig& cfg, int num_sms, int* tile_counter,
                              int max_depth, cudaStream_t stream) {
    if (in.n_tiles == 0) return cudaSuccess;
    const bool want_q = cfg.want_qsum || cfg.noise_model == 1;
    const bool coll = cfg.expect_collapsed != 0;
    cudaError_t e = cudaMemsetAsync(tile_counter, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    if (out.counts_out == nullptr && in.nib != nullptr && !want_q && !coll && cfg.tune_prefetch != 9) {
        // the hot path: PNIB16 (direction-split, nibble-packed); planes needed = bits of the largest (byte lane, allele) count = stored / 4
        const int need = in.nib_max_store / 4 + 2;
        // two sub-tiles per draw in both modes: the single-tile instance (--tune-prefetch 5, VCF mode only) is 10 % faster when the memory system is in its
        // fast state (0.170 against 0.187 ms) but waits on the tile-counter atomic twice as often and falls to 0.256 ms in the slow state this pool's
        // boxes are often in (profiles/r1_summary.md, "machine state"); the paired instance measures 0.187 ms in that state
        const bool pair = cfg.output_gvcf != 0 || cfg.tune_prefetch != 5;
        const int grid = max(1, min(num_sms * 4, (in.n_nib_tiles + (pair ? 15 : 7)) / (pair ? 16 : 8)));
        if (need < (1 << 8)) {
            if (pair) pileup_nib_score_kernel<8, true><<<grid, 256, 0, stream>>>(in, ex, out, cfg, tile_counter);
            else pileup_nib_score_kernel<8, false><<<grid, 256, 0, stream>>>(in, ex, out, cfg, tile_counter);
        } else {
            if (pair) pileup_nib_score_kernel<12, true><<<grid, 256, 0, stream>>>(in, ex, out, cfg, tile_counter);
            else pileup_nib_score_kernel<12, false><<<grid, 256, 0, stream>>>(in, ex, out, cfg, tile_counter);
        }
    } else if (out.counts_out == nullptr && max_depth + 2 * kChunk < (1 << 16)) {
        // vertical counters over the PTILE32 planes; planes needed = bits of the largest row count (entries + PADs of a locus)
        const int need = max_depth + 2 * kChunk;
        const int ctas = cfg.tune_ctas_per_sm == 3 ? 3 : 4;
        const int grid = max(1, min(num_sms * ctas, (in.n_tiles + 7) / 8));
#define PB2_VLAUNCH2(NP, Q, C)                                                                                                     \
    do {                                                                                                                           \
        if (NP == 10 && !Q && !C && cfg.tune_prefetch == 2) pileup_vcount_score_kernel<10, false, false, 4, 2><<<grid, 256, 0, stream>>>(in, ex, out, cfg, tile_counter); \
        else if (NP == 10 && !Q && !C && cfg.tune_prefetch == 3) pileup_vcount_score_kernel<10, false, false, 4, 3><<<grid, 256, 0, stream>>>(in, ex, out, cfg, tile_counter); \
        else if (NP == 10 && !Q && !C && cfg.tune_prefetch == 4) pileup_vcount_score_kernel<10, false, false, 4, 4><<<grid, 256, 0, stream>>>(in, ex, out, cfg, tile_counter); \
        else if (ctas == 3) pileup_vcount_score_kernel<NP, Q, C, 3><<<grid, 256, 0, stream>>>(in, ex, out, cfg, tile_counter);     \
        else pileup_vcount_score_kernel<NP, Q, C, 4><<<grid, 256, 0, stream>>>(in, ex, out, cfg, tile_counter);                    \
    } while (0)
#define PB2_VLAUNCH(NP)                                   \
    do {                                                  \
        if (want_q && coll) PB2_VLAUNCH2(NP, true, true); \
        else if (want_q) PB2_VLAUNCH2(NP, true, false);   \
        else if (coll) PB2_VLAUNCH2(NP, false, true);     \
        else PB2_VLAUNCH2(NP, false, false);              \
    } while (0)
        if (need < (1 << 10)) PB2_VLAUNCH(10);
        else if (need < (1 << 12)) PB2_VLAUNCH(12);
        else PB2_VLAUNCH(16);
#undef PB2_VLAUNCH
#undef PB2_VLAUNCH2
    } else {
        // general 198-bin variant (16-bit shared-memory histograms): pb2_get_counts, or loci deeper than 65 k entries are rejected by the caller
        const size_t smem = hot_kernel_smem_bytes(false, coll);
        const int grid = min(num_sms, (in.n_tiles + kHotThreads / 32 - 1) / (kHotThreads / 32));
#define PB2_LAUNCH(Q, C)                                                                                                              \
    do {                                                                                                                              \
        e = cudaFuncSetAttribute(pileup_count_score_kernel<uint16_t, kHotThreads, Q, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        if (e != cudaSuccess) return e;                                                                                               \
        pileup_count_score_kernel<uint16_t, kHotThreads, Q, C><<<grid, kHotThreads, smem, stream>>>(in, ex, out, cfg, tile_counter);  \
    } while (0)
        if (want_q && coll) PB2_LAUNCH(true, true);
        else if (want_q) PB2_LAUNCH(true, false);
        else if (coll) PB2_LAUNCH(false, true);
        else PB2_LAUNCH(false, false);
#undef PB2_LAUNCH
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    score_pending_kernel<<<num_sms * 4, 128, 0, stream>>>(in, ex, out, cfg);
    return cudaGetLastError();
}

// ================================================================================================ PVERT: read-major, bit-sliced pileup (pb2_pvert.cuh)
// The hot kernel of the reads path. Lane = locus, warp = tile of 32 loci. A block holds 32 rows (read stretches) of the tile bit-sliced per locus: two
// coalesced 16-byte loads give a lane the planes B0 B1 Q5..Q0 of its 32 slots. The quality rule `q < minBQ -> N` (RegionStateManager.cs:180-181) is a
// bit-serial comparison of the six quality planes against the bits of minBQ (one LOP3 per plane), the four allele indicators are one LOP3 each, and a
// count is a POPC: about 30 instructions per 32 entries, no counters to read out. Direction, collapsed-read category and entry kind are properties of
// the block's class, so stitched and collapsed-read data run through the same loop at the same byte per entry.
// kRefStream: the run has a dense reference stream (gVCF) - the per-locus tail then holds the inlined reference-allele scorer and the kernel keeps the
// 64-register / 4-CTA shape; without it the kernel runs 6 CTAs per SM. kAhead: blocks requested ahead of the one being counted.
template <bool kCollapsed, bool kRefStream, int kCtas, int kAhead>
__global__ void __launch_bounds__(256, kCtas)
pileup_pvert_score_kernel(const __grid_constant__ PvertPileup pv, const __grid_constant__ HotInputsExtra ex, const __grid_constant__ HotOutputs out,
                          const __grid_constant__ DeviceConfig cfg, int* __restrict__ tile_counter) {
    __shared__ __align__(16) PendingLocus s_pend[kCtaPending];
    __shared__ QueueScratch s_scratch;
    __shared__ int s_pend_n;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_pend_n = 0;
    __syncthreads();
    TilePileup in;   // what finish_locus / score_queued_locus read of the staged pileup
    in.ref_base = pv.ref_base; in.positions = pv.positions; in.first_position = pv.first_position; in.n_loci = pv.n_loci; in.n_tiles = pv.n_tiles;
    // bit i of minBQ as a lane-wide mask
    uint32_t M[6];
#pragma unroll
    for (int i = 0; i < 6; i++) M[i] = ((cfg.min_bq >> i) & 1) ? 0xffffffffu : 0u;
    const int nc = pv.n_classes;

    auto grab = [&]() -> int {
        int t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    int next_tile = grab();
    while (true) {
        const int tile = next_tile;
        if (tile >= pv.n_tiles) break;
        const uint8_t* p = pv.data + pv.tile_row0[tile] * 32 + lane * 16;
        const int32_t* ce = pv.cls_end + (int64_t)tile * nc;
        const int total = ce[nc - 1];   // rows of the tile (multiple of 32)
        int cnt[kNumAlleles][kNumDirs];
#pragma unroll
        for (int a = 0; a < kNumAlleles; a++)
#pragma unroll
            for (int d = 0; d < kNumDirs; d++) cnt[a][d] = 0;
        int coll[kNumCollapsed];
#pragma unroll
        for (int t = 0; t < kNumCollapsed; t++) coll[t] = 0;

        // blocks of all classes sit back to back: the loads run kAhead blocks ahead of the counting, across class boundaries (a ring of registers)
        uint4 rx[kAhead], ry[kAhead];
#pragma unroll
        for (int j = 0; j < kAhead; j++) {
            rx[j] = make_uint4(0, 0, 0, 0); ry[j] = rx[j];
            if (32 * j < total) { rx[j] = ldg_stream(p + 1024 * j); ry[j] = ldg_stream(p + 1024 * j + 512); }
        }
        int row = 0;
#pragma unroll 1
        for (int c = 0; c < nc; c++) {
            const int end = ce[c];
            if (end == row) continue;
            int a0 = 0, a1 = 0, a2 = 0, a3 = 0, pres = 0;
#pragma unroll 1
            for (; row < end; row += 32) {
                const uint4 x = rx[0], y = ry[0];
#pragma unroll
                for (int j = 0; j + 1 < kAhead; j++) { rx[j] = rx[j + 1]; ry[j] = ry[j + 1]; }
                if (row + 32 * kAhead < total) { rx[kAhead - 1] = ldg_stream(p + 1024 * kAhead); ry[kAhead - 1] = ldg_stream(p + 1024 * kAhead + 512); }
                p += 1024;
                // q >= minBQ, least significant plane first: ge_i = m_i ? (Q_i & ge) : (Q_i | ge)
                uint32_t ge = y.w | ~M[0];
                ge = (y.z & ge) | (~M[1] & (y.z | ge));
                ge = (y.y & ge) | (~M[2] & (y.y | ge));
                ge = (y.x & ge) | (~M[3] & (y.x | ge));
                ge = (x.w & ge) | (~M[4] & (x.w | ge));
                ge = (x.z & ge) | (~M[5] & (x.z | ge));
                a0 += __popc(ge & ~x.y & ~x.x);
                a1 += __popc(ge & ~x.y & x.x);
                a2 += __popc(ge & x.y & ~x.x);
                a3 += __popc(ge & x.y & x.x);
                pres += __popc(x.x | x.y | x.z | x.w | y.x | y.y | y.z | y.w);
            }
            const int kind = pv_class_kind(c), dir = pv_class_dir(c);
            const int counted = a0 + a1 + a2 + a3;
#pragma unroll
            for (int d = 0; d < kNumDirs; d++) {
                if (d != dir) continue;
                if (kind == 0) { cnt[AT_A][d] += a0; cnt[AT_G][d] += a1; cnt[AT_C][d] += a2; cnt[AT_T][d] += a3; cnt[AT_N][d] += pres - counted; }
                else cnt[AT_DEL][d] += counted;
            }
            if (kCollapsed) {   // CollapsedRegionState.AddCollapsedReadCount (:28-44): every entry not counted as N, typed reads only
                const int ct = pv_collapsed_code(pv_class_cg(c), dir);
#pragma unroll
                for (int t = 0; t < kNumCollapsed; t++) {
                    if (ct == t + 1) coll[t] += counted;
                    if (t == 2 && (ct - 1 == 4 || ct - 1 == 6)) coll[t] += counted;
                    if (t == 3 && (ct - 1 == 5 || ct - 1 == 7)) coll[t] += counted;
                }
            }
        }
        next_tile = grab();
        const int64_t locus = (int64_t)tile * kTileLoci + lane;
        if (locus >= pv.n_loci) continue;
        if (kCollapsed && out.collapsed_out != nullptr) {
#pragma unroll
            for (int t = 0; t < kNumCollapsed; t++) out.collapsed_out[locus * kNumCollapsed + t] = coll[t];
        }
        int any = 0;
#pragma unroll
        for (int a = 0; a < kNumAlleles; a++)
#pragma unroll
            for (int d = 0; d < kNumDirs; d++) any += cnt[a][d];
        const int ref_allele = allele_of_base(pv.ref_base[locus]);
        finish_locus<kRefStream>(cnt, 0.0, any, locus, ref_allele, in, ex, out, cfg, s_pend, &s_pend_n, kCtaTaskLoci);
    }

    __syncthreads();
    score_cta_queue(s_pend, min(s_pend_n, kCtaTaskLoci), s_scratch, in, ex, out, cfg);
}

cudaError_t launch_pvert_hot_kernel(const PvertPileup& pv, const HotInputsExtra& ex, const HotOutputs& out, const DeviceConfig& cfg, int num_sms, int* tile_counter,
                                    cudaStream_t stream) {
    if (pv.n_tiles == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(tile_counter, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    // the tuning runs of round 2 (profiles/r2_summary.md): 4 CTAs x 2 blocks ahead beat 2, 3, 5 and 6 CTAs per SM, 3 and 4 blocks ahead, and an L2 prefetch of
    // the next tile's head, for both instances
    const int grid = max(1, min(num_sms * 4, (pv.n_tiles + 7) / 8));
    if (out.ref_records != nullptr) {
        if (cfg.expect_collapsed) pileup_pvert_score_kernel<true, true, 4, 2><<<grid, 256, 0, stream>>>(pv, ex, out, cfg, tile_counter);
        else pileup_pvert_score_kernel<false, true, 4, 2><<<grid, 256, 0, stream>>>(pv, ex, out, cfg, tile_counter);
    } else {
        if (cfg.expect_collapsed) pileup_pvert_score_kernel<true, false, 4, 2><<<grid, 256, 0, stream>>>(pv, ex, out, cfg, tile_counter);
        else pileup_pvert_score_kernel<false, false, 4, 2><<<grid, 256, 0, stream>>>(pv, ex, out, cfg, tile_counter);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    TilePileup in;
    memset(&in, 0, sizeof(in));
    in.ref_base = pv.ref_base; in.positions = pv.positions; in.first_position = pv.first_position; in.n_loci = pv.n_loci; in.n_tiles = pv.n_tiles;
    score_pending_kernel<<<num_sms * 4, 128, 0, stream>>>(in, ex, out, cfg);
    return cudaGetLastError();
}

}  // namespace pb2
