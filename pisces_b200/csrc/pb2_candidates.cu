// Explicit-candidate kernels (sm_100a): the 198-bin count gather for the loci that spanning alleles touch, and the candidate scorer.
//
// Replaces (reference @ /root/reference):
//   IAlleleSource.GetAlleleCount / GetSumOfAlleleBaseQualities with anchor selection     src/lib/Pisces.Processing/RegionState/AlleleCountHelper.cs:21-85,102-166
//   CoverageCalculator.CalculateSpanning / CalculateSinglePoint / RedistributeStitchedCoverage   src/lib/Pisces.Calculators/CoverageCalculator.cs:49-98,162-331
//   AlleleCaller.ProcessVariant / IsCallable                                            src/exe/Pisces/Logic/VariantCalling/AlleleCaller.cs:208-258
//   AlleleProcessor.ApplyFilters incl. ComputeIndelRepeatLength                         AlleleProcessor.cs:25-213
//   RMxNCalculator.ComputeComponentRMxNLengths (all allele types)                       src/lib/Pisces.Calculators/RMxNCalculator.cs:49-133
//   SomaticGenotyper / SomaticGenotypeQualityCalculator (per allele)                    src/lib/Pisces.Genotyping/Somatic/*.cs
#include "pb2_candidates.cuh"
#include "pb2_math.cuh"

namespace pb2 {

__device__ __forceinline__ int cand_allele_of_base(uint8_t c) {  // AlleleHelper.GetAlleleType (Utility/AlleleHelper.cs:13-32)
    switch (c) { case 'A': return AT_A; case 'C': return AT_C; case 'G': return AT_G; case 'T': return AT_T; default: return AT_N; }
}

// ------------------------------------------------------------------------------------------------ per-locus count gather
// One warp per requested locus. The staged pileup interleaves the 16-entry chunks of the 32 loci of a tile step by step, ragged ends compacted,
// so the byte offset of (locus, step) is the tile base plus 16 x (active lanes of all earlier steps + active lanes below this one at this step):
// the warp replays the tile's activity masks from the 32 depths (one ballot per step), lane s % 32 keeps the offset of step s, and every 32
// steps all lanes load their chunk and add its entries to the warp's shared-memory histogram.
constexpr int kGatherWarps = 4;

__global__ void __launch_bounds__(32 * kGatherWarps)
gather_locus_counts_kernel(TilePileup in, const int32_t* __restrict__ req_locus, int32_t n_req, int32_t* __restrict__ out_counts,
                           int32_t* __restrict__ out_collapsed, double* __restrict__ out_qsum, int min_bq) {
    __shared__ int s_hist[kGatherWarps][kNumBins + kNumCollapsed];
    __shared__ double s_q[kGatherWarps][kNumBins];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * kGatherWarps + warp;
    if (r >= n_req) return;
    int* hist = s_hist[warp];
    double* qs = s_q[warp];
    for (int b = lane; b < kNumBins + kNumCollapsed; b += 32) hist[b] = 0;
    if (out_qsum != nullptr) for (int b = lane; b < kNumBins; b += 32) qs[b] = 0.0;
    __syncwarp();
    const int64_t locus = req_locus[r];
    if (locus >= 0 && locus < in.n_loci) {
        const int64_t tile = locus / kTileLoci;
        const int tl = (int)(locus % kTileLoci);
        const int64_t my_locus = tile * kTileLoci + lane;
        const int d = my_locus < in.n_loci ? in.depth[my_locus] : 0;
        const int nch = (d + kChunk - 1) / kChunk;
        const int n_target = __shfl_sync(0xffffffffu, nch, tl);
        const unsigned below = (1u << tl) - 1u;
        int64_t base = in.tile_base[tile];
        int64_t my_off = -1, my_cq = 0;
        int my_run = 0;
        for (int s = 0; s < n_target; s++) {
            const unsigned m = __ballot_sync(0xffffffffu, nch > s);
            if ((s & 31) == lane) { my_off = base + (int64_t)__popc(m & below) * kChunk; my_cq = base + my_off; my_run = __popc(m) * kChunk; }
            base += (int64_t)__popc(m) * kChunk;
            if ((s & 31) == 31 || s == n_target - 1) {
                if (my_off >= 0) {
                    const uint4 wc = *reinterpret_cast<const uint4*>(in.cq + my_cq);             // code chunks of the step, then its quality chunks
                    const uint4 wq = *reinterpret_cast<const uint4*>(in.cq + my_cq + my_run);
                    const uint4 wa = *reinterpret_cast<const uint4*>(in.anch + my_off);
                    const uint32_t cw[4] = {wc.x, wc.y, wc.z, wc.w}, qw[4] = {wq.x, wq.y, wq.z, wq.w}, aw[4] = {wa.x, wa.y, wa.z, wa.w};
#pragma unroll
                    for (int k = 0; k < kChunk; k++) {
                        const uint32_t code = (cw[k >> 2] >> ((k & 3) * 8)) & 0xffu, q = (qw[k >> 2] >> ((k & 3) * 8)) & 0x7fu, an = (aw[k >> 2] >> ((k & 3) * 8)) & 0xffu;   // staged q carries bit 7
                        int allele = (int)(code & 7u);
                        if (allele == 7 || (int)q < min_bq) allele = AT_N;    // staged N is 7; RegionStateManager.cs:180-181: low quality -> N
                        const int dir = (int)((code >> 3) & 3u);
                        const int bin = (allele * kNumDirs + dir) * kNumAnchors + (int)(an & 15u);
                        atomicAdd(&hist[bin], 1);
                        if (allele != AT_N) {
                            const int ct = (int)(an >> 4);
                            if (ct != 0) {   // CollapsedRegionState.AddCollapsedReadCount (:28-44)
                                atomicAdd(&hist[kNumBins + ct - 1], 1);
                                if (ct - 1 == 4 || ct - 1 == 6) atomicAdd(&hist[kNumBins + 2], 1);
                                else if (ct - 1 == 5 || ct - 1 == 7) atomicAdd(&hist[kNumBins + 3], 1);
                            }
                            if (out_qsum != nullptr && allele != AT_DEL) atomicAdd(&qs[bin], pow(10.0, (double)((float)(-(int)q) / 10.0f)));   // :191 float exponent
                        }
                    }
                }
                my_off = -1;
                __syncwarp();
            }
        }
        __syncwarp();
        if (lane == 0) hist[(AT_N * kNumDirs + DIR_F) * kNumAnchors + 0] -= in.pad[locus];   // PAD entries of the staged locus
        __syncwarp();
    }
    for (int b = lane; b < kNumBins; b += 32) out_counts[(int64_t)r * kNumBins + b] = hist[b];
    if (out_collapsed != nullptr && lane < kNumCollapsed) out_collapsed[(int64_t)r * kNumCollapsed + lane] = hist[kNumBins + lane];
    if (out_qsum != nullptr) for (int b = lane; b < kNumBins; b += 32) out_qsum[(int64_t)r * kNumBins + b] = qs[b];
}

cudaError_t launch_gather_locus_counts(const TilePileup& in, const int32_t* req_locus, int32_t n_req, int32_t* out_counts, int32_t* out_collapsed,
                                       double* out_qsum, int min_bq, cudaStream_t stream) {
    if (n_req <= 0) return cudaSuccess;
    const int blocks = (n_req + kGatherWarps - 1) / kGatherWarps;
    gather_locus_counts_kernel<<<blocks, 32 * kGatherWarps, 0, stream>>>(in, req_locus, n_req, out_counts, out_collapsed, out_qsum, min_bq);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ anchor selection
// AlleleCountHelper.GetAnchorAdjustedAlleleCount / ...TotalQuality (AlleleCountHelper.cs:21-85,102-166) with symmetric = false, K = 5.
// max_anchor < 0 = null.
template <class T>
__device__ __forceinline__ T anchor_adjusted(const T* __restrict__ bins, int min_anchor, int max_anchor, bool from_end) {
    const int true_min = min(kAnchorK, min_anchor);
    int initial_max = kAnchorK;
    if (max_anchor >= 0) initial_max = max_anchor >= kAnchorK ? kAnchorK - 1 : max_anchor;
    T tot = 0;
    if (from_end) {
        for (int i = true_min; i <= initial_max; i++) tot += bins[kNumAnchors - i - 1];
        if (max_anchor < 0) for (int i = 0; i < initial_max; i++) tot += bins[i];
    } else {
        for (int i = true_min; i <= initial_max; i++) tot += bins[i];
        if (max_anchor < 0) for (int i = initial_max + 1; i < kNumAnchors; i++) tot += bins[i];
    }
    return tot;
}

// ------------------------------------------------------------------------------------------------ repeats
__device__ __forceinline__ bool bytes_equal(const uint8_t* a, const uint8_t* b, int n) {
    for (int k = 0; k < n; k++) if (a[k] != b[k]) return false;
    return true;
}
// One bookend of RMxNCalculator.ComputeRMxNLengthForIndel (:49-95): the prefix (pass 0) or suffix (pass 1) of the variant bases starting at offset i,
// walked back over the reference while it repeats, then counted forward. Counts are capped at `cap` (callers compare min/max of them against cap only;
// going back further than cap units cannot change a count that is capped at cap).
__device__ int cand_rmxn_bookend(int variant_position, const uint8_t* vb, int length, int i, int pass, const uint8_t* __restrict__ ref, int64_t ref_len, int cap) {
    const int blen = length - i;
    const uint8_t* book = pass == 0 ? vb : vb + i;
    int64_t back = variant_position;
    for (int steps = 0; steps < cap; steps++) {
        const int64_t nb = back - blen;
        if (nb < 0 || nb + blen > ref_len) break;
        if (!bytes_equal(ref + nb, book, blen)) break;
        back = nb;
    }
    int reps = 0;
    int64_t cur = back;
    while (reps < cap) {
        if (cur < 0 || cur + blen > ref_len) break;
        if (!bytes_equal(ref + cur, book, blen)) break;
        reps++;
        cur += blen;
    }
    return reps;
}
// RMxNCalculator.ShouldFilter (:19-38,104-133) for any allele type, by a full warp: the (up to 3 calls x 2 x max unit length) bookend scans are
// independent chains of dependent reference reads, so each lane takes one and the per-call maxima are warp-reduced.
__device__ bool cand_rmxn_should_filter_warp(const DevCand& c, const uint8_t* ref_allele, const uint8_t* alt_allele, const DeviceConfig& cfg,
                                             const uint8_t* __restrict__ chr, int64_t chr_len) {
    const int lane = threadIdx.x & 31;
    const int cap = max(cfg.rmxn_min_reps, 1);
    const int U = max(cfg.rmxn_max_len, 1);
    int n_calls;
    int vpos[3], vlen[3];
    const uint8_t* vb[3];
    if (c.type == CAT_INS) { n_calls = 1; vpos[0] = c.position; vb[0] = alt_allele + 1; vlen[0] = c.alt_len - 1; }
    else if (c.type == CAT_DEL) { n_calls = 1; vpos[0] = c.position; vb[0] = ref_allele + 1; vlen[0] = c.ref_len - 1; }
    else {
        n_calls = 3;
        vpos[0] = c.position - 1; vb[0] = ref_allele; vlen[0] = c.ref_len;
        vpos[1] = c.position + c.ref_len - 1; vb[1] = alt_allele; vlen[1] = c.alt_len;
        vpos[2] = c.position - 1; vb[2] = alt_allele; vlen[2] = c.alt_len;
    }
    int best[3] = {0, 0, 0};
    for (int t = lane; t < n_calls * 2 * U; t += 32) {
        const int call = t / (2 * U), pass = (t / U) & 1, ui = t % U;
        const int units = min(cfg.rmxn_max_len, vlen[call]);
        if (ui >= units) continue;
        const int i = vlen[call] - units + ui;
        best[call] = max(best[call], cand_rmxn_bookend(vpos[call], vb[call], vlen[call], i, pass, chr, chr_len, cap));
    }
    for (int k = 0; k < 3; k++) best[k] = __reduce_max_sync(0xffffffffu, best[k]);
    const int c1 = best[0];
    const int c2 = n_calls == 3 ? max(best[1], best[2]) : INT32_MAX;
    return min(c1, c2) >= cfg.rmxn_min_reps;
}

// AlleleProcessor.ComputeIndelRepeatLength (:80-135) with SimplifyRepeatUnit (:140-155) and GetRepeatLength (:160-213). `bases` of the reference
// (upstream + downstream, 50 flanking bases each) is the contiguous chromosome window [ub, de].
__device__ int cand_indel_repeat_length(const DevCand& c, const uint8_t* ref_allele, const uint8_t* alt_allele, const uint8_t* __restrict__ chr, int64_t chr_len) {
    if (chr == nullptr || chr_len == 0) return 0;
    if (c.type != CAT_INS && c.type != CAT_DEL && c.type != CAT_SNV) return 0;
    const int64_t sp = (int64_t)c.position - 1;
    int64_t ub = sp - 50, ue = sp - 1, db = sp, de = sp + 50 - 1;
    if (ub < 0) ub = 0;
    if (db < 0) db = 0;
    if (de >= chr_len) de = chr_len - 1;
    if (ue >= chr_len) ue = chr_len - 1;
    const int64_t up_len = ue >= 0 ? ue - ub + 1 : 0;
    const int64_t down_len = de - db + 1;   // Substring throws in the reference when negative; positions are inside the chromosome here
    if (down_len < 0) return 0;
    const int64_t total = up_len + down_len;
    const uint8_t* bases = chr + (up_len > 0 ? ub : db);
    int current = (int)up_len;
    const uint8_t* vb = nullptr;
    int vlen = 0;
    if (c.type == CAT_INS) { vb = alt_allele + 1; vlen = c.alt_len - 1; current++; }
    if (c.type == CAT_DEL) { vb = ref_allele + 1; vlen = c.ref_len - 1; current++; }
    if (vlen == 0) return 0;   // GetRepeatLength: empty repeat unit
    // SimplifyRepeatUnit: shortest prefix whose non-overlapping occurrences (scanning left to right) cover the unit's length
    int ulen = 1;
    for (int i = 1; i < vlen; i++) {
        int occ = 0, pos = 0;
        while (pos + ulen <= vlen) {
            if (bytes_equal(vb + pos, vb, ulen)) { occ++; pos += ulen; } else pos++;
        }
        if (vlen == occ * ulen) break;
        ulen++;
    }
    // GetRepeatLength
    const int last_position = (int)total - ulen - 1;
    if (current + ulen + 1 > (int)total) return 1;
    int previous = current;
    while (current > 0) {
        if (!bytes_equal(bases + current, vb, ulen)) break;
        previous = current;
        current -= ulen;
    }
    current = previous;
    int repeat = 0;
    while (current <= last_position) {
        if (!bytes_equal(bases + current, vb, ulen)) break;
        current += ulen;
        repeat++;
    }
    return repeat;
}

// ------------------------------------------------------------------------------------------------ the candidate scorer
// One warp per candidate: the lanes stage the two 198-bin count rows in shared memory (coalesced) and share the RMxN reference scans; lane 0 does
// the coverage arithmetic, the q-score / strand-bias / genotype chain and writes the record.
constexpr int kScoreWarps = 4;

__global__ void __launch_bounds__(32 * kScoreWarps) score_candidates_kernel(CandScoreArgs a, DeviceConfig cfg) {
    __shared__ int32_t s_cnt[kScoreWarps][2][kNumBins];
    __shared__ double s_qs[kScoreWarps][2][kNumBins];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * kScoreWarps + warp;
    if (i >= a.n) return;
    const DevCand c = a.cands[i];
    const uint8_t* ref_allele = a.arena + c.allele_off;
    const uint8_t* alt_allele = ref_allele + c.ref_len;
    const bool want_q = a.qsum != nullptr;
    for (int b = lane; b < kNumBins; b += 32) {
        s_cnt[warp][0][b] = c.req_start >= 0 ? a.counts[(int64_t)c.req_start * kNumBins + b] : 0;
        s_cnt[warp][1][b] = c.req_end >= 0 ? a.counts[(int64_t)c.req_end * kNumBins + b] : 0;
        if (want_q) {
            s_qs[warp][0][b] = c.req_start >= 0 ? a.qsum[(int64_t)c.req_start * kNumBins + b] : 0.0;
            s_qs[warp][1][b] = c.req_end >= 0 ? a.qsum[(int64_t)c.req_end * kNumBins + b] : 0.0;
        }
    }
    __syncwarp();
    int allele_support = c.support[0] + c.support[1] + c.support[2];
    if (c.type == CAT_REF) allele_support = max(0, allele_support - c.gapped_ref);   // CoverageCalculator.cs:94-97
    const bool is_ref = c.type == CAT_REF;
    // the reference scans need no coverage: all lanes, before lane 0 goes on alone (skipped by ShouldFilter when the frequency is above the limit; the
    // limit test is applied to the result below)
    bool rmxn_hit = false;
    if (!is_ref && cfg.rmxn_max_len >= 0 && cfg.rmxn_min_reps >= 0 && a.chr_seq != nullptr) rmxn_hit = cand_rmxn_should_filter_warp(c, ref_allele, alt_allele, cfg, a.chr_seq, a.chr_len);
    if (lane != 0) return;

    int cov[3] = {0, 0, 0};
    int total = 0, ref_support = 0, nocalls = 0;
    double qsum = 0.0;
    const int cover_alleles[5] = {AT_A, AT_C, AT_G, AT_T, AT_DEL};   // Constants.CoverageContributingAlleles (order matters for the double sums)

    if (c.type == CAT_SNV || c.type == CAT_REF) {
        // CalculateSinglePoint (:49-98)
        const int32_t* cs = s_cnt[warp][0];
        const double* qs = want_q ? s_qs[warp][0] : nullptr;
        const int ref_type = c.ref_len == 1 ? cand_allele_of_base(ref_allele[0]) : AT_N;
        for (int d = 0; d < 3; d++) {
            for (int k = 0; k < 5; k++) {
                const int at = cover_alleles[k];
                const int n = cs ? anchor_adjusted<int>(cs + (at * kNumDirs + d) * kNumAnchors, 0, -1, false) : 0;
                cov[d] += n;
                if (qs) qsum += anchor_adjusted<double>(qs + (at * kNumDirs + d) * kNumAnchors, 0, -1, false);
                if (at == ref_type) ref_support += n;
            }
            total += cov[d];
            nocalls += cs ? anchor_adjusted<int>(cs + (AT_N * kNumDirs + d) * kNumAnchors, 0, -1, false) : 0;
        }
        if (c.type == CAT_SNV) ref_support = max(0, ref_support - c.gapped_ref);
    } else {
        // CalculateSpanning (:162-321)
        const int length = c.type == CAT_DEL ? c.ref_len - 1 : c.type == CAT_INS ? c.alt_len - 1 : c.alt_len;   // BaseAllele.Length (:24-43)
        const bool presume_anchored = c.type == CAT_INS ? (cfg.expect_stitched != 0) : true;
        const int32_t* cs = s_cnt[warp][0];
        const int32_t* ce = s_cnt[warp][1];
        const double* qs = want_q ? s_qs[warp][0] : nullptr;
        const double* qe = want_q ? s_qs[warp][1] : nullptr;
        const bool picky = c.type == CAT_INS;   // considerAnchorInformation: TrackedAnchorSize > 0 (Factory.cs:193-199)
        int first_base = AT_N, last_base = AT_N;
        if (picky) { first_base = cand_allele_of_base(alt_allele[1]); last_base = cand_allele_of_base(alt_allele[c.alt_len - 1]); }
        SpanIngredients g;
        for (int d = 0; d < 3; d++) { g.sp[d] = 0; g.ep[d] = 0; g.spu[d] = 0; g.epu[d] = 0; }
        g.conf_l = 0; g.conf_r = 0; g.susp_l = 0; g.susp_r = 0;
        double unq_start = 0, unq_end = 0;
        const int unanchored_support = allele_support - c.well_anchored;
        for (int d = 0; d < 3; d++) {
            for (int k = 0; k < 5; k++) {
                const int at = cover_alleles[k];
                const int off = (at * kNumDirs + d) * kNumAnchors;
                const int min_end = (picky && at == first_base) ? length : 0;
                const int min_start = (picky && at == last_base) ? length : 0;
                const int s = cs ? anchor_adjusted<int>(cs + off, min_start, -1, false) : 0;
                const int e = ce ? anchor_adjusted<int>(ce + off, min_end, -1, true) : 0;
                g.sp[d] += s; g.ep[d] += e; g.conf_l += s; g.conf_r += e;
                if (qs) qsum += anchor_adjusted<double>(qs + off, min_start, -1, false);
                if (qe) qsum += anchor_adjusted<double>(qe + off, min_end, -1, true);
                if (picky) {   // collected unconditionally; spanning_tail applies the `unanchoredSupport > 0` condition of :222
                    if (min_start > 0) {
                        const int n = cs ? anchor_adjusted<int>(cs + off, 0, min_start - 1, false) : 0;
                        g.spu[d] += n; g.susp_l += n;
                        if (qs && unanchored_support > 0) unq_start += anchor_adjusted<double>(qs + off, 0, min_start - 1, false);
                    }
                    if (min_end > 0) {
                        const int n = ce ? anchor_adjusted<int>(ce + off, 0, min_end - 1, true) : 0;
                        g.epu[d] += n; g.susp_r += n;
                        if (qs && unanchored_support > 0) unq_end += anchor_adjusted<double>(qs + off, 0, min_end - 1, true);   // the reference reads the START position here (:253)
                    }
                }
            }
        }
        const SpanCoverage sc = spanning_tail(g, picky, presume_anchored, allele_support, c.well_anchored);
        if (picky) for (int d = 0; d < 3; d++) { qsum += unq_start * (double)sc.weight; qsum += unq_end * (double)sc.weight; }
        for (int d = 0; d < 3; d++) cov[d] = sc.cov[d];
        total = sc.total;
        if (a.out_ingredients != nullptr) a.out_ingredients[i] = g;
        ref_support = max(0, total - allele_support);
    }

    // ---- AlleleCaller.ProcessVariant (:208-234)
    const float freq = allele_frequency(allele_support, total);
    int vq = 0, nl_applied = 0;
    SbResult sb;
    sb.bias = 0; sb.gatk = 0; sb.acceptable = false; sb.var_both = false; sb.cov_both = false;
    if (allele_support > 0) {
        int nl = cfg.noise_level;
        double error_rate = cfg.vq_error_rate;
        if (cfg.noise_model == 1) {
            nl = (int)(-10 * log10(qsum / total));
            error_rate = q_to_p((double)nl);
        }
        nl_applied = nl;
        vq = (total == 0) ? 0 : poisson_qscore(allele_support, total, error_rate, cfg.max_vq);
        sb = strand_bias(cov, c.support, cfg.sb_noise, (double)cfg.sb_acceptance, cfg.sb_model, cfg.sb_min_vf);
    }
    const float all_reads = (float)(total + nocalls);
    const float frac_nc = all_reads == 0 ? 0.0f : ((float)nocalls / all_reads);
    unsigned filters = 0;
    if (cfg.low_depth_filter >= 0 && total < cfg.low_depth_filter) filters |= 1u << FLT_LOW_DEPTH;
    if (vq < cfg.vq_filter && total != 0) filters |= 1u << FLT_LOW_VQ;
    if (!is_ref) {
        if (cfg.no_call_filter >= 0 && frac_nc > cfg.no_call_filter) filters |= 1u << FLT_NO_CALL;
        if (!sb.acceptable || (cfg.filter_single_strand && !sb.var_both)) filters |= 1u << FLT_STRAND_BIAS;
        if (a.indel_repeat_filter > 0 && a.indel_repeat_filter <= cand_indel_repeat_length(c, ref_allele, alt_allele, a.chr_seq, a.chr_len)) filters |= 1u << FLT_INDEL_REPEAT;
        if (rmxn_hit && !(freq >= cfg.rmxn_freq_limit)) filters |= 1u << FLT_RMXN;
        if (freq < cfg.variant_freq_filter) filters |= 1u << FLT_LOW_VF;
        if (cfg.expect_stitched && (c.flags & kCandAltHasN)) filters |= 1u << FLT_STRAND_BIAS;
    }
    const float ref_freq = allele_frequency(ref_support, total);
    // AlleleCaller.IsCallable (:236-258) && ShouldReport (:260-263)
    bool callable = true;
    if (!is_ref) {
        if (total < cfg.min_coverage && !cfg.output_gvcf) callable = false;
        if (total != 0 && freq < cfg.min_frequency) callable = false;
        if (vq < cfg.min_vq) callable = false;
    }
    const bool report = callable && (c.flags & kCandReportable);
    // a forced allele that would not be reported is reported anyway, flagged, and keeps the genotype a new CalledAllele starts with (:108-118,150)
    const bool forced_report = (c.flags & kCandForced) && !report;
    if (forced_report) filters |= 1u << FLT_FORCED_REPORT;
    int gt = is_ref ? GT_HOM_REF : GT_HET_ALT_REF, gq = 0;
    if (forced_report) {
    } else if (cfg.ploidy == PLOIDY_SOMATIC) {
        gt = somatic_genotype(is_ref, total, freq, ref_freq, cfg.min_frequency_filter, cfg.min_coverage);
        gq = somatic_gq(gt, vq, total, freq, cfg.target_lod, cfg.min_gq, cfg.max_gq, a.q_to_p_table, a.q_table_max);
    } else if (is_ref) {   // germline: exact for a reference-only locus; loci with variants are genotyped together in pb2_flush
        const bool hap = cfg.ploidy == PLOIDY_HAPLOID;
        gt = germline_reference_only_genotype(hap, total, allele_support, ref_support, cfg.diploid_minor_vf, cfg.diploid_major_vf, cfg.min_coverage);
        gq = germline_gq(hap, gt, total, allele_support, cfg.min_gq, cfg.max_gq);
    }
    if (cfg.low_gq_filter >= 0 && (float)gq < (float)cfg.low_gq_filter) filters |= 1u << FLT_LOW_GQ;

    pb2_call_record r;
    r.position = c.position;
    r.type = c.type;
    r.genotype = (uint8_t)gt;
    r.sb_flags = (sb.acceptable ? 1 : 0) | (sb.var_both ? 2 : 0) | (sb.cov_both ? 4 : 0) | (forced_report ? 8 : 0);
    r.open_flags = 0;
    r.filters = (uint16_t)filters;
    r.noise_level = (uint16_t)nl_applied;
    r.variant_qscore = vq;
    r.genotype_qscore = gq;
    r.total_coverage = total;
    for (int d = 0; d < 3; d++) { r.coverage_by_direction[d] = cov[d]; r.support_by_direction[d] = c.support[d]; }
    r.allele_support = allele_support;
    r.reference_support = ref_support;
    r.num_no_calls = nocalls;
    r.fraction_no_calls = frac_nc;
    if (c.ref_len + c.alt_len <= 4) {
        uint32_t ab = 0;
        for (int k = 0; k < c.ref_len + c.alt_len; k++) ab |= (uint32_t)ref_allele[k] << (8 * k);
        r.allele_bytes = ab;
    } else r.allele_bytes = c.allele_off;
    r.ref_len = (uint16_t)c.ref_len;
    r.alt_len = (uint16_t)c.alt_len;
    r.sum_base_quality = qsum;
    r.bias_score = sb.bias;
    r.gatk_bias_score = sb.gatk;

    if (a.out_dense != nullptr) a.out_dense[i] = r;
    if (a.out_callable != nullptr) a.out_callable[i] = (callable ? 1 : 0) | (report ? 2 : 0) | (forced_report ? 4 : 0);
    if (a.var_records != nullptr && report && (cfg.own_hi <= 0 || (c.position >= cfg.own_lo && c.position <= cfg.own_hi))) {
        const unsigned long long slot = atomicAdd(a.var_count, 1ull);
        if ((int64_t)slot < a.var_capacity) a.var_records[slot] = r;
        if (a.ref_valid != nullptr && c.locus >= 0 && !is_ref) a.ref_valid[c.locus] = 0;
    }
}

// After a scorer pass that ran next to the hot kernel: prune the reference records of the positions where an explicit allele was called
// (AlleleCaller.cs:146-147); the hot kernel has written ref_valid by now.
__global__ void prune_ref_valid_kernel(const DevCand* __restrict__ cands, const uint8_t* __restrict__ flags, int32_t n, uint8_t* __restrict__ ref_valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (flags[i] & 2) && cands[i].locus >= 0 && cands[i].type != CAT_REF) ref_valid[cands[i].locus] = 0;
}
cudaError_t launch_prune_ref_valid(const DevCand* cands, const uint8_t* flags, int32_t n, uint8_t* ref_valid, cudaStream_t stream) {
    if (n <= 0 || ref_valid == nullptr) return cudaSuccess;
    prune_ref_valid_kernel<<<(n + 255) / 256, 256, 0, stream>>>(cands, flags, n, ref_valid);
    return cudaGetLastError();
}

cudaError_t launch_score_candidates(const CandScoreArgs& args, const DeviceConfig& cfg, cudaStream_t stream) {
    if (args.n <= 0) return cudaSuccess;
    score_candidates_kernel<<<(args.n + kScoreWarps - 1) / kScoreWarps, 32 * kScoreWarps, 0, stream>>>(args, cfg);
    return cudaGetLastError();
}

}  // namespace pb2
