// C ABI of libpisces_b200.so (include/pisces_b200.h): handle, staging, launches. No CPU compute path exists here: every
// entry point that produces counts or calls runs the CUDA kernels in pb2_kernels.cu and fails with PB2_ERR_CUDA otherwise.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include <thread>
#include "pb2_internal.hpp"

using namespace pb2;

// PB2_TRACE=1: wall-clock breakdown of pushes and flushes on stderr (diagnostics only)
#include <chrono>
static bool trace_on() { static const bool on = getenv("PB2_TRACE") != nullptr; return on; }
struct Trace {
    const char* what;
    std::chrono::steady_clock::time_point t0, last;
    std::string line;
    explicit Trace(const char* w) : what(w), t0(std::chrono::steady_clock::now()), last(t0) {}
    void mark(const char* label) {
        if (!trace_on()) return;
        const auto now = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof buf, " %s=%.2fms", label, std::chrono::duration<double, std::milli>(now - last).count());
        line += buf;
        last = now;
    }
    ~Trace() {
        if (!trace_on()) return;
        fprintf(stderr, "[pb2] %s total=%.2fms%s\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), line.c_str());
    }
};

#include <nvtx3/nvToolsExt.h>
struct nvtx_range {   // NVTX ranges around push / stage / resident step / flush (tracing hook)
    explicit nvtx_range(const char* name) { nvtxRangePushA(name); }
    ~nvtx_range() { nvtxRangePop(); }
};

#define CU(h, expr)                                                                                      \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            (h)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                             \
            return PB2_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

static thread_local std::string g_create_error;
int pb2_fail(pb2_handle* h, int code, const std::string& msg) {
    if (h) h->error = msg; else g_create_error = msg;
    return code;
}
static int fail(pb2_handle* h, int code, const std::string& msg) { return pb2_fail(h, code, msg); }
// SmallVariantCaller's constructor (SmallVariantCaller.cs:48-77): the forced alleles of this chromosome, by position, minus those outside the intervals
static void rearm_forced(pb2_handle* h);

extern "C" void pb2_default_config(pb2_config* c) {
    memset(c, 0, sizeof(*c));
    c->device = 0;
    c->min_base_call_quality = 20;      // BamFilterParameters.cs:8
    c->min_frequency = 0.01f;           // VariantCallingParameters.cs:59
    c->min_frequency_filter = -1;
    c->target_lod_frequency = -1;
    c->max_variant_qscore = 100; c->min_variant_qscore = 20; c->variant_qscore_filter = 30;
    c->max_genotype_qscore = 100; c->min_genotype_qscore = 0; c->low_genotype_quality_filter = -1;
    c->min_coverage = 10; c->low_depth_filter = -1;
    c->rmxn_max_repeat_len = 5; c->rmxn_min_repetitions = 9; c->rmxn_frequency_limit = 0.35f;
    c->forced_noise_level = -1; c->noise_model = 0;
    c->strand_bias_acceptance = 0.5f; c->strand_bias_model = 1; c->filter_single_strand = 0;
    c->no_call_filter = 0.6f; c->ploidy = 0; c->tracked_anchor_size = 5; c->output_gvcf = 1;
    c->expect_stitched = 0; c->expect_collapsed = 0; c->want_sum_base_quality = 0; c->collapse = 1; c->call_mnvs = 0;
    c->indel_repeat_filter = -1;        // VariantCallingParameters.cs:74 (null)
    c->max_size_mnv = 3; c->max_gap_mnv = 1;   // PiscesApplicationOptions.cs:43-66
    c->collapse_freq_threshold = 0.0f; c->collapse_freq_ratio_threshold = 0.5f; c->exclude_mnvs_from_collapsing = 0;
    c->diploid_minor_vf = 0.20f; c->diploid_major_vf = 0.70f; c->diploid_sum_vf_multiallelic = 0.80f;   // VariantCallingParameters.cs:84
    c->is_male = -1;
    c->amplicon_bias_filter = -1.0f;
}

extern "C" int pb2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* pb2_last_error(pb2_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

static void derive_config(pb2_handle* h) {
    const pb2_config& c = h->cfg;
    DeviceConfig& d = h->dcfg;
    d.min_bq = c.min_base_call_quality;
    d.noise_level = c.forced_noise_level == -1 ? c.min_base_call_quality : c.forced_noise_level;  // VariantCallingParameters.cs:109-118
    d.noise_model = c.noise_model;
    // GenotypeCreator.GetPloidyForThisChr (GenotypeCreator.cs:39-68)
    {
        const std::string& n = h->chr_name;
        int p = c.ploidy;
        if (p == PLOIDY_SOMATIC || n == "chrM" || n == "M") p = PLOIDY_SOMATIC;
        else if (p == PLOIDY_HAPLOID) p = PLOIDY_HAPLOID;
        else if (c.is_male < 0) {}
        else if (c.is_male > 0 && (n == "chrY" || n == "chrX" || n == "Y" || n == "X")) p = PLOIDY_HAPLOID;
        else if (c.is_male == 0 && (n == "chrY" || n == "Y")) p = PLOIDY_HAPLOID;
        d.ploidy = p;
    }
    d.diploid_minor_vf = c.diploid_minor_vf; d.diploid_major_vf = c.diploid_major_vf; d.diploid_sum_vf = c.diploid_sum_vf_multiallelic;
    // VariantCallerConfig.MinFrequency = genotypeCalculator.MinVarFrequency (Factory.cs:160): MinimumFrequency for the somatic genotyper, MinorVF otherwise
    d.min_frequency = d.ploidy == PLOIDY_SOMATIC ? c.min_frequency : c.diploid_minor_vf;
    d.sb_min_vf = (double)d.min_frequency;
    d.min_frequency_filter = c.min_frequency_filter < c.min_frequency ? c.min_frequency : c.min_frequency_filter;  // Validate :144-147
    d.target_lod = c.target_lod_frequency < d.min_frequency_filter ? d.min_frequency_filter : c.target_lod_frequency;  // :152-155
    if (c.skip_validation) { d.min_frequency_filter = c.min_frequency_filter; d.target_lod = c.target_lod_frequency; }
    d.variant_freq_filter = d.min_frequency_filter > d.min_frequency ? d.min_frequency_filter : d.min_frequency;  // IGenotypeCalculator.SetMinFreqFilter
    d.max_vq = c.max_variant_qscore; d.min_vq = c.min_variant_qscore; d.vq_filter = c.variant_qscore_filter;
    d.max_gq = c.max_genotype_qscore; d.min_gq = c.min_genotype_qscore; d.low_gq_filter = c.low_genotype_quality_filter;
    d.min_coverage = c.min_coverage;
    d.low_depth_filter = c.low_depth_filter < c.min_coverage ? c.min_coverage : c.low_depth_filter;  // :137-141
    if (c.skip_validation) d.low_depth_filter = c.low_depth_filter;   // -1 = null: no LowDepth filter
    d.rmxn_max_len = c.rmxn_max_repeat_len; d.rmxn_min_reps = c.rmxn_min_repetitions; d.rmxn_freq_limit = c.rmxn_frequency_limit;
    d.sb_acceptance = c.strand_bias_acceptance; d.sb_model = c.strand_bias_model; d.filter_single_strand = c.filter_single_strand;
    d.no_call_filter = c.no_call_filter;
    d.output_gvcf = c.output_gvcf; d.expect_stitched = c.expect_stitched; d.expect_collapsed = c.expect_collapsed;
    d.have_intervals = h->have_intervals ? 1 : 0;
    d.want_qsum = c.want_sum_base_quality;
    d.one = 1;
    d.tune_prefetch = c.reserved[1];
    d.tune_ctas_per_sm = c.reserved[0] == 3 ? 3 : 4;   // tuning knob (bench.py --tune-ctas)
    d.snv_from_counts = c.call_mnvs ? 0 : 1;   // CallMNVs: SNV candidates come from the finder's state machine (explicit), not from the counts
    d.own_lo = h->own_lo; d.own_hi = h->own_hi;
    d.vq_error_rate = std::pow(10.0, -1 * (double)d.noise_level / 10.0);                      // QtoP: double division (MathOperations.cs:7-10)
    d.sb_noise = std::pow(10.0, (double)((float)(-1 * d.noise_level) / 10.0f));               // float exponent (StrandBiasCalculator.cs:32)
}

extern "C" int pb2_create(const pb2_config* cfg, pb2_handle** out) {
    if (!cfg || !out) return fail(nullptr, PB2_ERR_ARG, "pb2_create: null argument");
    *out = nullptr;
    if (cfg->ploidy != PLOIDY_SOMATIC && cfg->ploidy != PLOIDY_DIPLOID && cfg->ploidy != PLOIDY_HAPLOID)
        return fail(nullptr, PB2_ERR_UNSUPPORTED, "pb2_create: ploidy must be 0 Somatic, 1 DiploidByThresholding or 3 Haploid (DiploidByAdaptiveGT is not built, SURVEY 8f rank 4)");
    if (cfg->tracked_anchor_size != 5) return fail(nullptr, PB2_ERR_UNSUPPORTED, "pb2_create: tracked_anchor_size must be 5");
    if (cfg->min_base_call_quality < 0 || cfg->min_base_call_quality > 127) return fail(nullptr, PB2_ERR_ARG, "pb2_create: min_base_call_quality must be in [0,127]");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, PB2_ERR_CUDA, std::string("pb2_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
    }
    if (cfg->device < 0 || cfg->device >= n) return fail(nullptr, PB2_ERR_ARG, "pb2_create: bad device ordinal");
    pb2_handle* h = new pb2_handle();
    h->cfg = *cfg;
    h->device = cfg->device;
    derive_config(h);
    auto bail = [&](const char* what, cudaError_t err) { g_create_error = std::string(what) + ": " + cudaGetErrorString(err); delete h; return PB2_ERR_CUDA; };
    if ((e = cudaSetDevice(h->device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, h->device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    h->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&h->ev_stage0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&h->ev_stage1)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (int i = 0; i < 2; i++) {
        if ((e = cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&h->ev_scattered[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&h->ev_piece[2 * i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&h->ev_piece[2 * i + 1], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    }
    {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = h->device;
        if ((e = cudaMemPoolCreate(&h->pool, &props)) != cudaSuccess) return bail("cudaMemPoolCreate", e);
        uint64_t keep = UINT64_MAX;
        if ((e = cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return bail("cudaMemPoolSetAttribute", e);
    }
    if ((e = cudaMalloc(&h->d_tile_counter, sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaHostAlloc(&h->h_counters, sizeof(unsigned long long) * 5, cudaHostAllocDefault)) != cudaSuccess) return bail("cudaHostAlloc", e);
    {
        h->q_table_max = std::min(std::max(cfg->max_variant_qscore, 0), 1023);
        std::vector<double> t((size_t)h->q_table_max + 1);
        for (int q = 0; q <= h->q_table_max; q++) t[(size_t)q] = std::pow(10.0, -1 * (double)q / 10.0);   // MathOperations.QtoP
        if ((e = cudaMalloc(&h->d_q_to_p, t.size() * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
        if ((e = cudaMemcpy(h->d_q_to_p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy", e);
    }
    if (h->dcfg.ploidy == PLOIDY_SOMATIC) {
        if ((e = cudaMalloc(&h->d_gq_tail, (sizeof(double) + sizeof(int)) * kGqTailMaxCov * kGqTailMaxA)) != cudaSuccess) return bail("cudaMalloc", e);
        // the finished-GQ half of the table is for a variant q-score at its cap; p1 is the very double the kernels would read from d_q_to_p
        h->gq_capped_vq = (h->dcfg.max_vq >= 0 && h->dcfg.max_vq <= h->q_table_max) ? h->dcfg.max_vq : -1;
        const double p1 = h->gq_capped_vq >= 0 ? std::pow(10.0, -1 * (double)h->gq_capped_vq / 10.0) : 0.0;
        if ((e = launch_gq_tail_fill(h->d_gq_tail, h->dcfg.target_lod, p1, h->dcfg.min_gq, h->dcfg.max_gq, h->stream)) != cudaSuccess) return bail("gq_tail_fill", e);
        if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return bail("gq_tail_fill", e);
    }
    *out = h;
    return PB2_OK;
}

// Device memory of a handle comes from its own stream-ordered pool (cudaMallocFromPoolAsync on the handle's stream) that never returns memory to
// the driver while the handle lives: after the first push every staging buffer is a pool hit, so a push costs no cudaMalloc / cudaFree.
// On top of the pool sits a small block cache: a freed block is kept by the handle and handed out again to the next request of about its size (within
// 25 %), so a host that repeats the same push / flush cycle makes no allocator call at all in steady state (growing the pool by gigabytes costs tens to
// hundreds of milliseconds and showed up as a x3 spread of the end-to-end step). All users are ordered on the handle's stream (the copy / side stream
// work is joined by events before a block is freed), so a block is reused in stream order like a pool allocation.
cudaError_t pb2_dev_alloc(pb2_handle* h, void** p, size_t bytes) {
    bytes = (std::max<size_t>(bytes, 256) + 255) & ~(size_t)255;
    long best = -1;
    for (size_t i = 0; i < h->blocks.size(); i++) {
        const DevBlock& b = h->blocks[i];
        if (!b.used && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 + 4096 && (best < 0 || b.bytes < h->blocks[(size_t)best].bytes)) best = (long)i;
    }
    if (best >= 0) { h->blocks[(size_t)best].used = true; *p = h->blocks[(size_t)best].p; h->cached_free_bytes -= h->blocks[(size_t)best].bytes; return cudaSuccess; }
    const cudaError_t e = cudaMallocFromPoolAsync(p, bytes, h->pool, h->stream);
    if (e == cudaSuccess) h->blocks.push_back(DevBlock{*p, bytes, true});
    return e;
}
void pb2_dev_free(pb2_handle* h, void* p) {
    if (!p) return;
    for (size_t i = 0; i < h->blocks.size(); i++)
        if (h->blocks[i].p == p) {
            h->blocks[i].used = false;
            h->cached_free_bytes += h->blocks[i].bytes;
            // bound what the cache holds back: beyond 64 idle blocks or 24 GB the oldest idle blocks return to the pool
            size_t idle = 0;
            for (auto& b : h->blocks) idle += b.used ? 0 : 1;
            for (size_t k = 0; k < h->blocks.size() && (idle > 64 || h->cached_free_bytes > ((size_t)24 << 30));) {
                if (!h->blocks[k].used) { cudaFreeAsync(h->blocks[k].p, h->stream); h->cached_free_bytes -= h->blocks[k].bytes; h->blocks.erase(h->blocks.begin() + (long)k); idle--; }
                else k++;
            }
            return;
        }
    cudaFreeAsync(p, h->stream);   // not one of ours (allocated before the cache existed)
}
static void release_block_cache(pb2_handle* h) {
    for (auto& b : h->blocks) if (!b.used) cudaFreeAsync(b.p, h->stream);
    h->blocks.erase(std::remove_if(h->blocks.begin(), h->blocks.end(), [](const DevBlock& b) { return !b.used; }), h->blocks.end());
    h->cached_free_bytes = 0;
}
static cudaError_t pool_alloc(pb2_handle* h, void** p, size_t bytes) { return pb2_dev_alloc(h, p, bytes); }
static void pool_free(pb2_handle* h, void* p, size_t = 0) { pb2_dev_free(h, p); }
template <class T>
static cudaError_t pool_alloc_t(pb2_handle* h, T** p, size_t count) { return pool_alloc(h, reinterpret_cast<void**>(p), count * sizeof(T)); }

static void release_resident_graph(pb2_handle* h) {
    if (h->resident_graph) { cudaGraphExecDestroy(h->resident_graph); h->resident_graph = nullptr; }
    h->resident_graph_failed = false;
}

static void free_segment(pb2_handle* h, Segment& s) {
    void* ptrs[] = {s.code, s.anch, s.ref_records, s.var_records, s.pending, s.depth, s.pad, s.tile_base, s.ref_base, s.positions, s.ref_valid, s.exc_entries, s.counters,
                    s.nib, s.nib_tile_base, s.nib_store, s.nib_depth, s.pv_data, s.pv_row_meta, s.pv_tile_row0, s.pv_cls_end, s.pv_row_amp};
    for (void* p : ptrs) pool_free(h, p);
    s = Segment();
}

static void free_reads(pb2_handle* h);
extern "C" int pb2_reset(pb2_handle* h) {
    if (!h) return PB2_ERR_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto& s : h->segs) free_segment(h, s);
    h->segs.clear();
    h->h_out.clear();
    h->h_out_ext.clear();
    free_reads(h);
    h->cands.clear(); h->cand_by_pos.clear(); h->block_max_endpoint.clear(); h->gapped_ref.clear(); h->triggers.clear(); h->arena.clear();
    h->last_trigger_key = 0; h->push_last_key = 0; h->cleared_through = 0;
    h->snv_explicit_ranges.clear();
    rearm_forced(h);
    explicit_release_resident(h);
    release_resident_graph(h);
    return PB2_OK;
}

extern "C" void pb2_destroy(pb2_handle* h) {
    if (!h) return;
    pb2_reset(h);
    if (h->d_chr) cudaFree(h->d_chr);
    if (h->d_tile_counter) cudaFree(h->d_tile_counter);
    if (h->h_counters) cudaFreeHost(h->h_counters);
    if (h->pin_refs) cudaFreeHost(h->pin_refs);
    if (h->pin_valid) cudaFreeHost(h->pin_valid);
    if (h->d_q_to_p) cudaFree(h->d_q_to_p);
    if (h->d_gq_tail) cudaFree(h->d_gq_tail);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_stage0) cudaEventDestroy(h->ev_stage0);
    if (h->ev_stage1) cudaEventDestroy(h->ev_stage1);
    for (int i = 0; i < 2; i++) { if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]); if (h->ev_scattered[i]) cudaEventDestroy(h->ev_scattered[i]); }
    for (cudaEvent_t ev : h->ev_piece) if (ev) cudaEventDestroy(ev);
    release_block_cache(h);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->pool) cudaMemPoolDestroy(h->pool);
    delete h;
}

extern "C" void* pb2_stream(pb2_handle* h) { return h ? (void*)h->stream : nullptr; }

extern "C" int pb2_set_reference(pb2_handle* h, const char* chr_name, const uint8_t* seq, int64_t len) {
    if (!h || !chr_name || (!seq && len > 0) || len < 0) return fail(h, PB2_ERR_ARG, "pb2_set_reference: bad argument");
    CU(h, cudaSetDevice(h->device));
    h->chr_name = chr_name;
    h->h_chr.assign(seq, seq + len);
    for (auto& b : h->h_chr) if (b >= 'a' && b <= 'z') b = (uint8_t)(b - 32);   // Genome.cs:81-96 upper-cases on load
    if (h->d_chr) { cudaFree(h->d_chr); h->d_chr = nullptr; }
    h->chr_len = len;
    if (len > 0) {
        CU(h, cudaMalloc(&h->d_chr, (size_t)len + 16));   // slack: kernels read the chromosome a word at a time
        CU(h, cudaMemsetAsync(h->d_chr + len, 'N', 16, h->stream));
        CU(h, cudaMemcpyAsync(h->d_chr, h->h_chr.data(), (size_t)len, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    }
    derive_config(h);   // the genotyper follows the chromosome (GenotypeCreator.GetPloidyForThisChr)
    explicit_release_resident(h);
    release_resident_graph(h);
    return PB2_OK;
}

extern "C" int pb2_set_intervals(pb2_handle* h, const int32_t* start, const int32_t* end, int32_t n) {
    if (!h || n < 0 || (n > 0 && (!start || !end))) return fail(h, PB2_ERR_ARG, "pb2_set_intervals: bad argument");
    h->iv_start.assign(start, start + n);
    h->iv_end.assign(end, end + n);
    h->have_intervals = true;   // an empty set still means "intervals applied" (Factory.cs:229-245)
    rearm_forced(h);
    derive_config(h);
    return PB2_OK;
}

static int push_common(pb2_handle* h, const pb2_pileup_csr* p, bool device_ptrs) {
    const bool packed = p && p->layout == PB2_LAYOUT_PACKED2;
    if (!h || !p || p->n_loci < 0 || (p->n_loci > 0 && (!p->offsets || !p->code || !p->qual || (!packed && !p->anchor))))
        return fail(h, PB2_ERR_ARG, "pb2_push_pileup: bad argument");
    if (p->layout != PB2_LAYOUT_PLANES && !packed) return fail(h, PB2_ERR_ARG, "pb2_push_pileup: unknown layout");
    if (packed && (p->n_flags < 0 || (p->n_flags > 0 && (!p->flag_index || !p->flag_bits)))) return fail(h, PB2_ERR_ARG, "pb2_push_pileup: bad flag list");
    if (packed && h->cfg.expect_collapsed) return fail(h, PB2_ERR_ARG, "pb2_push_pileup: PB2_LAYOUT_PACKED2 has no room for the collapsed-read type (expect_collapsed)");
    if (p->n_loci == 0) return PB2_OK;
    if (p->n_loci > (int64_t)1 << 31) return fail(h, PB2_ERR_ARG, "pb2_push_pileup: more than 2^31 loci in one push");
    if (!p->ref_bases && !p->positions && (h->chr_len == 0)) return fail(h, PB2_ERR_STATE, "pb2_push_pileup: no ref_bases given and no reference set");
    CU(h, cudaSetDevice(h->device));
    release_resident_graph(h);
    Trace tr("push_pileup");
    cudaStream_t st = h->stream;
    Segment s;
    s.n_loci = p->n_loci;
    s.n_tiles = (int32_t)((p->n_loci + kTileLoci - 1) / kTileLoci);
    s.first_position = p->first_position;
    s.has_positions = p->positions != nullptr;

    // CSR offsets on the device (the entry planes follow in chunks below)
    const int64_t* d_off = nullptr;
    int64_t* tmp_off = nullptr;
    int64_t n_entries = 0;
    if (device_ptrs) {
        d_off = p->offsets;
        CU(h, cudaMemcpyAsync(&n_entries, p->offsets + p->n_loci, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
    } else {
        n_entries = p->offsets[p->n_loci];
        if (n_entries < 0) return fail(h, PB2_ERR_ARG, "pb2_push_pileup: negative entry count");
        CU(h, pool_alloc_t(h, &tmp_off, (size_t)(p->n_loci + 1)));
        CU(h, cudaMemcpyAsync(tmp_off, p->offsets, sizeof(int64_t) * (size_t)(p->n_loci + 1), cudaMemcpyHostToDevice, st));
        d_off = tmp_off;
    }
    s.n_entries = n_entries;

    // per-locus side arrays
    CU(h, pool_alloc_t(h, &s.depth, (size_t)p->n_loci));
    CU(h, pool_alloc_t(h, &s.pad, (size_t)p->n_loci));
    CU(h, pool_alloc_t(h, &s.tile_base, (size_t)(s.n_tiles + 1)));
    CU(h, pool_alloc_t(h, &s.ref_base, (size_t)p->n_loci));
    const cudaMemcpyKind kind = device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (p->positions) {
        CU(h, pool_alloc_t(h, &s.positions, (size_t)p->n_loci));
        CU(h, cudaMemcpyAsync(s.positions, p->positions, sizeof(int32_t) * (size_t)p->n_loci, kind, st));
        s.h_positions.resize((size_t)p->n_loci);
        CU(h, cudaMemcpyAsync(s.h_positions.data(), p->positions, sizeof(int32_t) * (size_t)p->n_loci, device_ptrs ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost, st));
    }
    if (p->ref_bases) {
        CU(h, cudaMemcpyAsync(s.ref_base, p->ref_bases, (size_t)p->n_loci, kind, st));
    } else {
        // contiguous positions: slice of the chromosome already on the device
        if (p->positions) return fail(h, PB2_ERR_ARG, "pb2_push_pileup: ref_bases is required together with positions");
        const int64_t a = (int64_t)p->first_position - 1;
        if (a < 0 || a + p->n_loci > h->chr_len) return fail(h, PB2_ERR_ARG, "pb2_push_pileup: loci outside the reference");
        CU(h, cudaMemcpyAsync(s.ref_base, h->d_chr + a, (size_t)p->n_loci, cudaMemcpyDeviceToDevice, st));
    }

    // layout: depths, per-tile sizes, exclusive scan -> tile_base
    int64_t* tile_bytes = nullptr;
    CU(h, pool_alloc_t(h, &tile_bytes, (size_t)(s.n_tiles + 1)));
    CU(h, cudaMemsetAsync(tile_bytes, 0, sizeof(int64_t) * (size_t)(s.n_tiles + 1), st));
    int32_t* d_max_depth = reinterpret_cast<int32_t*>(h->d_tile_counter);   // scratch int of the handle (the hot kernel resets it before use)
    CU(h, cudaMemsetAsync(d_max_depth, 0, sizeof(int32_t), st));
    CU(h, launch_tile_layout(d_off, p->n_loci, s.depth, tile_bytes, d_max_depth, st));
    CU(h, cudaMemcpyAsync(&s.max_depth, d_max_depth, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    size_t temp_bytes = 0;
    CU(h, exclusive_scan_i64(tile_bytes, s.tile_base, s.n_tiles + 1, nullptr, 0, &temp_bytes, st));
    void* temp = nullptr;
    CU(h, pool_alloc(h, &temp, std::max<size_t>(temp_bytes, 16)));
    CU(h, exclusive_scan_i64(tile_bytes, s.tile_base, s.n_tiles + 1, temp, temp_bytes, nullptr, st));
    CU(h, cudaMemcpyAsync(&s.plane_bytes, s.tile_base + s.n_tiles, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    tr.mark("layout");
    if (s.max_depth >= 65000) return fail(h, PB2_ERR_UNSUPPORTED, "pb2_push_pileup: a locus with 65000 or more entries (16-bit counters)");
    const size_t pb = (size_t)std::max<int64_t>(s.plane_bytes, 16) + 4096;   // slack: the hot kernel prefetches up to two steps past a tile
    s.alloc_plane = pb;
    CU(h, pool_alloc(h, (void**)&s.code, 2 * pb));   // code + quality plane
    CU(h, pool_alloc(h, (void**)&s.anch, pb));
    // the side list of flagged entries is filled by the scatter below (counters[3]; counters[0..2] are reset by every run)
    s.exc_capacity = 1 << 20;
    CU(h, pool_alloc_t(h, &s.exc_entries, 2 * (size_t)s.exc_capacity));
    CU(h, pool_alloc_t(h, &s.counters, 5));   // variants, -, pending, flagged entries, amplicon status
    CU(h, cudaMemsetAsync(s.counters, 0, sizeof(unsigned long long) * 5, st));
    h->total_launches += 3;

    // the interleaving scatter
    // sparse candidate flags of the packed layout, on the device
    int64_t* d_flag_index = nullptr;
    uint8_t* d_flag_bits = nullptr;
    if (packed && p->n_flags > 0 && !device_ptrs) {
        CU(h, pool_alloc_t(h, &d_flag_index, (size_t)p->n_flags));
        CU(h, pool_alloc_t(h, &d_flag_bits, (size_t)p->n_flags));
        CU(h, cudaMemcpyAsync(d_flag_index, p->flag_index, sizeof(int64_t) * (size_t)p->n_flags, cudaMemcpyHostToDevice, st));
        CU(h, cudaMemcpyAsync(d_flag_bits, p->flag_bits, (size_t)p->n_flags, cudaMemcpyHostToDevice, st));
    }
    if (device_ptrs) {
        const uint8_t *pc = p->code, *pq = p->qual, *pa = p->anchor;
        uint8_t* tmp[3] = {nullptr, nullptr, nullptr};
        if (packed) {   // unpack into scratch planes (the caller's buffers stay as they are)
            for (int k = 0; k < 3; k++) CU(h, pool_alloc(h, (void**)&tmp[k], (size_t)std::max<int64_t>(n_entries, 16)));
            CU(h, cudaMemcpyAsync(tmp[0], p->code, (size_t)n_entries, cudaMemcpyDeviceToDevice, st));
            CU(h, cudaMemcpyAsync(tmp[1], p->qual, (size_t)n_entries, cudaMemcpyDeviceToDevice, st));
            CU(h, launch_unpack_packed2(tmp[0], tmp[1], tmp[2], n_entries, st));
            CU(h, launch_apply_entry_flags(tmp[0], p->flag_index, p->flag_bits, p->n_flags, 0, n_entries, st));
            h->total_launches += 2;
            pc = tmp[0]; pq = tmp[1]; pa = tmp[2];
        }
        CU(h, launch_tile_scatter(d_off, pc, pq, pa, p->n_loci, 0, s.n_tiles, 0, s.tile_base, s.ref_base, h->dcfg.min_bq, s.code, s.anch, s.pad, s.exc_entries,
                                  s.counters + 3, s.exc_capacity, st));
        h->total_launches += 1;
        for (int k = 0; k < 3; k++) pool_free(h, tmp[k]);
    } else {
        // Host planes: chunks of whole tiles through two staging buffers; the H2D copy of chunk i+1 (copy stream) overlaps the scatter of chunk i
        // (handle stream), so a push costs the PCIe time of the three planes and little else.
        const int64_t chunk_target = (int64_t)48 << 20;   // entries (= bytes per plane) per chunk
        int64_t max_chunk = 0;
        std::vector<std::pair<int32_t, int32_t>> chunks;   // [tile0, tile1)
        for (int32_t t0 = 0; t0 < s.n_tiles;) {
            const int64_t e0 = p->offsets[(int64_t)t0 * kTileLoci];
            int32_t lo = t0 + 1, hi = s.n_tiles;           // first tile end whose entry count reaches the target (offsets are monotone)
            while (lo < hi) {
                const int32_t mid = lo + (hi - lo) / 2;
                if (p->offsets[std::min<int64_t>((int64_t)mid * kTileLoci, p->n_loci)] - e0 >= chunk_target) hi = mid; else lo = mid + 1;
            }
            const int32_t t1 = lo;
            max_chunk = std::max(max_chunk, p->offsets[std::min<int64_t>((int64_t)t1 * kTileLoci, p->n_loci)] - e0);
            chunks.push_back({t0, t1});
            t0 = t1;
        }
        uint8_t* stage[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
        const int n_buf = chunks.size() > 1 ? 2 : 1;
        for (int b = 0; b < n_buf; b++) for (int k = 0; k < 3; k++) CU(h, pool_alloc(h, (void**)&stage[b][k], (size_t)std::max<int64_t>(max_chunk, 16)));
        CU(h, cudaEventRecord(h->ev_scattered[0], st));   // the copy stream must not run ahead of the allocations made on the handle stream
        CU(h, cudaStreamWaitEvent(h->copy_stream, h->ev_scattered[0], 0));
        for (size_t i = 0; i < chunks.size(); i++) {
            const int b = (int)(i % 2);
            const int32_t t0 = chunks[i].first, t1 = chunks[i].second;
            const int64_t e0 = p->offsets[(int64_t)t0 * kTileLoci], e1 = p->offsets[std::min<int64_t>((int64_t)t1 * kTileLoci, p->n_loci)];
            if (i >= 2) CU(h, cudaStreamWaitEvent(h->copy_stream, h->ev_scattered[b], 0));
            if (e1 > e0) {
                CU(h, cudaMemcpyAsync(stage[b][0], p->code + e0, (size_t)(e1 - e0), cudaMemcpyHostToDevice, h->copy_stream));
                CU(h, cudaMemcpyAsync(stage[b][1], p->qual + e0, (size_t)(e1 - e0), cudaMemcpyHostToDevice, h->copy_stream));
                if (!packed) CU(h, cudaMemcpyAsync(stage[b][2], p->anchor + e0, (size_t)(e1 - e0), cudaMemcpyHostToDevice, h->copy_stream));
            }
            CU(h, cudaEventRecord(h->ev_copied[b], h->copy_stream));
            CU(h, cudaStreamWaitEvent(st, h->ev_copied[b], 0));
            if (packed && e1 > e0) {
                CU(h, launch_unpack_packed2(stage[b][0], stage[b][1], stage[b][2], e1 - e0, st));
                CU(h, launch_apply_entry_flags(stage[b][0], d_flag_index, d_flag_bits, p->n_flags, e0, e1, st));
                h->total_launches += 2;
            }
            CU(h, launch_tile_scatter(d_off, stage[b][0], stage[b][1], stage[b][2], p->n_loci, t0, t1 - t0, e0, s.tile_base, s.ref_base, h->dcfg.min_bq, s.code, s.anch, s.pad,
                                      s.exc_entries, s.counters + 3, s.exc_capacity, st));
            CU(h, cudaEventRecord(h->ev_scattered[b], st));
            h->total_launches += 1;
        }
        for (int b = 0; b < n_buf; b++) for (int k = 0; k < 3; k++) pool_free(h, stage[b][k]);
    }

    tr.mark("enqueue_chunks");
    // ---- PNIB16: the direction-split, nibble-packed form the hot kernel reads (1.5 B / entry), derived from the staged planes on the device. Not for
    // collapsed-read tracking or quality sums (they need the third byte / per-entry work) and not when a Stitched-direction entry or a very deep locus
    // shows up: those segments stay with the PTILE32 kernel.
    if (!h->cfg.expect_collapsed && !h->cfg.want_sum_base_quality && h->cfg.noise_model != 1 && h->cfg.reserved[1] != 9) {
        s.n_nib_tiles = (int32_t)((p->n_loci + kNibLoci - 1) / kNibLoci);
        int64_t* nib_tile_bytes = nullptr;
        int32_t* d_flags = nullptr;
        CU(h, pool_alloc_t(h, &s.nib_store, 2 * (size_t)p->n_loci));
        CU(h, pool_alloc_t(h, &s.nib_depth, 2 * (size_t)p->n_loci));
        CU(h, pool_alloc_t(h, &s.nib_tile_base, (size_t)s.n_nib_tiles + 1));
        CU(h, pool_alloc_t(h, &nib_tile_bytes, (size_t)s.n_nib_tiles + 1));
        CU(h, pool_alloc_t(h, &d_flags, 2));
        CU(h, cudaMemsetAsync(nib_tile_bytes, 0, sizeof(int64_t) * ((size_t)s.n_nib_tiles + 1), st));
        CU(h, cudaMemsetAsync(d_flags, 0, sizeof(int32_t) * 2, st));
        TilePileup view;
        memset(&view, 0, sizeof(view));
        view.cq = s.code; view.anch = s.anch; view.tile_base = s.tile_base; view.depth = s.depth; view.pad = s.pad; view.ref_base = s.ref_base;
        view.n_loci = s.n_loci; view.n_tiles = s.n_tiles; view.n_nib_tiles = s.n_nib_tiles;
        CU(h, launch_nib_count(view, s.nib_store, s.nib_depth, nib_tile_bytes, d_flags, st));
        size_t nib_temp_bytes = 0;
        CU(h, exclusive_scan_i64(nib_tile_bytes, s.nib_tile_base, s.n_nib_tiles + 1, nullptr, 0, &nib_temp_bytes, st));
        void* nib_temp = nullptr;
        CU(h, pool_alloc(h, &nib_temp, std::max<size_t>(nib_temp_bytes, 16)));
        CU(h, exclusive_scan_i64(nib_tile_bytes, s.nib_tile_base, s.n_nib_tiles + 1, nib_temp, nib_temp_bytes, nullptr, st));
        int32_t flags[2] = {0, 0};
        CU(h, cudaMemcpyAsync(&s.nib_bytes, s.nib_tile_base + s.n_nib_tiles, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CU(h, cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
        h->total_launches += 3;
        if (flags[0] == 0 && flags[1] <= kChunk * kNibMaxChunks) {
            s.nib_max_store = flags[1];
            CU(h, pool_alloc(h, (void**)&s.nib, (size_t)s.nib_bytes + 4096));
            CU(h, launch_nib_scatter(view, s.nib_store, s.nib_tile_base, s.nib, st));
            h->total_launches += 1;
        } else {
            pool_free(h, s.nib_store); pool_free(h, s.nib_depth); pool_free(h, s.nib_tile_base);
            s.nib_store = s.nib_depth = nullptr; s.nib_tile_base = nullptr; s.n_nib_tiles = 0;
        }
        pool_free(h, nib_tile_bytes);
        pool_free(h, d_flags);
        pool_free(h, nib_temp);
    }
    tr.mark("pnib16");
    // outputs
    if (h->cfg.output_gvcf) {
        s.alloc_ref = sizeof(pb2_call_record) * (size_t)p->n_loci;
        CU(h, pool_alloc(h, (void**)&s.ref_records, s.alloc_ref));
        CU(h, pool_alloc_t(h, &s.ref_valid, (size_t)p->n_loci));
    }
    s.var_capacity = std::max<int64_t>(1024, p->n_loci);
    s.alloc_var = sizeof(pb2_call_record) * (size_t)s.var_capacity;
    CU(h, pool_alloc(h, (void**)&s.var_records, s.alloc_var));
    s.pending_capacity = std::max<int64_t>(1024, p->n_loci);
    s.alloc_pending = sizeof(PendingLocus) * (size_t)s.pending_capacity;
    CU(h, pool_alloc(h, (void**)&s.pending, s.alloc_pending));
    pool_free(h, tile_bytes);
    pool_free(h, temp);
    pool_free(h, tmp_off);
    pool_free(h, d_flag_index);
    pool_free(h, d_flag_bits);
    CU(h, cudaStreamSynchronize(st));   // the caller's buffers are consumed when this returns
    tr.mark("copy+scatter");
    h->segs.push_back(std::move(s));
    return PB2_OK;
}

extern "C" int pb2_push_pileup(pb2_handle* h, const pb2_pileup_csr* p) { return push_common(h, p, false); }
extern "C" int pb2_push_pileup_device(pb2_handle* h, const pb2_pileup_csr* p) { return push_common(h, p, true); }

static const char* const kTooManyAmplicons = "Index was outside the bounds of the array.";   // a seventh amplicon name at one position (RegionState.cs:293-297)
// AmpliconBiasCalculator.Compute over the segment's variant stream (count-based SNVs, and explicit ones appended by the resident explicit pass)
static int amplicon_pass(pb2_handle* h, Segment& s, cudaStream_t st) {
    if (s.pv_row_amp == nullptr || h->cfg.amplicon_bias_filter < 0) return PB2_OK;
    CU(h, cudaMemsetAsync(s.counters + 4, 0, sizeof(unsigned long long), st));
    CU(h, launch_pvert_amplicon_bias(pvert_view(s), s.pv_row_amp, s.var_records, s.counters, 0, s.var_capacity, nullptr, h->cfg.amplicon_bias_filter, h->dcfg.min_bq,
                                     reinterpret_cast<int*>(s.counters + 4), h->num_sms, st));
    h->total_launches += 1;
    return PB2_OK;
}
// enqueue: counters reset, hot kernel (+ overflow scorer) between the timing events. finish: counters back, one synchronize, bookkeeping.
static int enqueue_segment(pb2_handle* h, Segment& s, int32_t* counts_out, int32_t* collapsed_out, const int32_t* d_gapped, bool reset_counters = true,
                           bool capturing = false) {
    cudaStream_t st = h->stream;
    TilePileup in;
    memset(&in, 0, sizeof(in));
    in.cq = s.code; in.anch = s.anch; in.tile_base = s.tile_base; in.depth = s.depth; in.pad = s.pad; in.ref_base = s.ref_base;
    in.positions = s.positions; in.first_position = s.first_position; in.n_loci = s.n_loci; in.n_tiles = s.n_tiles; in.plane_bytes = std::max<int64_t>(s.plane_bytes, 16);
    in.nib = s.nib; in.nib_tile_base = s.nib_tile_base; in.nib_store = s.nib_store; in.nib_depth = s.nib_depth; in.n_nib_tiles = s.n_nib_tiles; in.nib_max_store = s.nib_max_store;
    HotInputsExtra ex;
    ex.gapped_ref = d_gapped; ex.locus_has_variant = nullptr; ex.chr_seq = h->d_chr; ex.chr_len = h->chr_len; ex.q_to_p_table = h->d_q_to_p; ex.q_table_max = h->q_table_max; ex.gq_tail_table = h->d_gq_tail; ex.gq_capped_vq = h->gq_capped_vq;
    HotOutputs out;
    out.ref_records = s.ref_records; out.ref_valid = s.ref_valid; out.var_records = s.var_records; out.var_count = s.counters;
    out.var_capacity = s.var_capacity; out.exc_entries = s.exc_entries; out.exc_count = s.counters + 3; out.exc_capacity = s.exc_capacity;
    out.counts_out = counts_out; out.collapsed_out = collapsed_out;
    out.pending = s.pending; out.pending_count = s.counters + 2; out.pending_capacity = s.pending_capacity;
    if (reset_counters) CU(h, cudaMemsetAsync(s.counters, 0, sizeof(unsigned long long) * 3, st));
    // inside a stream capture the timing events must be external event-record nodes to be readable with cudaEventElapsedTime
    CU(h, cudaEventRecordWithFlags(h->ev0, st, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
    if (s.pv_data != nullptr) CU(h, launch_pvert_hot_kernel(pvert_view(s), ex, out, h->dcfg, h->num_sms, h->d_tile_counter, st));
    else CU(h, launch_hot_kernel(in, ex, out, h->dcfg, h->num_sms, h->d_tile_counter, s.max_depth, st));
    CU(h, cudaEventRecordWithFlags(h->ev1, st, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
    return PB2_OK;
}
static int finish_segment(pb2_handle* h, Segment& s, bool counters_already_copied = false) {
    cudaStream_t st = h->stream;
    unsigned long long* cnt4 = h->h_counters;
    if (!counters_already_copied) CU(h, cudaMemcpyAsync(cnt4, s.counters, sizeof(unsigned long long) * 5, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    float ms = 0;
    CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->hot_ms += ms;
    h->hot_launches += 1;
    h->total_launches += 2;
    const unsigned long long cnt[2] = {cnt4[0], cnt4[3]};
    s.h_var_count = cnt[0];
    s.h_exc_count = cnt[1];
    s.called = true;
    if ((int64_t)cnt[0] > s.var_capacity) return fail(h, PB2_ERR_NOMEM, "variant record buffer overflow");
    if ((int64_t)cnt[1] > s.exc_capacity) return fail(h, PB2_ERR_NOMEM, "open-ended candidate side list overflow");
    if (s.pv_row_amp != nullptr && cnt4[4] != 0) return fail(h, PB2_ERR_ARG, kTooManyAmplicons);
    return PB2_OK;
}
static int run_segment(pb2_handle* h, Segment& s, int32_t* counts_out, int32_t* collapsed_out, const int32_t* d_gapped = nullptr) {
    int rc = enqueue_segment(h, s, counts_out, collapsed_out, d_gapped);
    if (rc == PB2_OK) rc = amplicon_pass(h, s, h->stream);
    if (rc == PB2_OK) rc = finish_segment(h, s);
    if (rc == PB2_ERR_NOMEM && (int64_t)h->h_counters[0] > s.var_capacity) {
        // up to three SNV alleles per locus can be callable (permissive thresholds on deep, noisy data): the stream is sized for one per locus, the
        // counter holds the exact need - grow and run the segment again
        const int64_t need = (int64_t)h->h_counters[0];
        pool_free(h, s.var_records);
        s.var_records = nullptr;
        s.var_capacity = need + 1024;
        s.alloc_var = sizeof(pb2_call_record) * (size_t)s.var_capacity;
        CU(h, pool_alloc(h, (void**)&s.var_records, s.alloc_var));
        rc = enqueue_segment(h, s, counts_out, collapsed_out, d_gapped);
        if (rc == PB2_OK) rc = amplicon_pass(h, s, h->stream);
        if (rc == PB2_OK) rc = finish_segment(h, s);
    }
    return rc;
}

// ------------------------------------------------------------------------------------------------ the device read store
template <class T>
static cudaError_t grow(pb2_handle* h, GrowBuf<T>& b, size_t need, size_t keep) {
    if (need <= b.cap) return cudaSuccess;
    const size_t ncap = std::max(need, b.cap + b.cap / 2 + 64);
    T* np = nullptr;
    cudaError_t e = pool_alloc_t(h, &np, ncap);
    if (e != cudaSuccess) return e;
    if (b.p && keep) e = cudaMemcpyAsync(np, b.p, keep * sizeof(T), cudaMemcpyDeviceToDevice, h->stream);
    pool_free(h, b.p);
    b.p = np; b.cap = ncap;
    return e;
}
// the reads leave the store, its buffers stay for the next push (pb2_reset / pb2_destroy release them)
static void clear_reads(pb2_handle* h) {
    DeviceReads& R = h->reads;
    R.n = R.n_cigar = R.n_seq = 0;
    R.has_dirs = R.has_collapsed = R.has_amplicon = false;
    R.min_start = INT32_MAX; R.max_end = 0; R.last_pos0 = -1;
}
static void free_reads(pb2_handle* h) {
    DeviceReads& R = h->reads;
    void* ptrs[] = {R.pos0.p, R.end_pos.p, R.flag.p, R.cigar_off.p, R.seq_off.p, R.cigar.p, R.bases.p, R.quals.p, R.base_dirs.p, R.collapsed.p, R.slots.p, R.amplicon.p};
    for (void* p : ptrs) pool_free(h, p);
    h->reads = DeviceReads();
}
// per-base directions of reads pushed without base_dirs: the read's own direction (Read.cs:390-421 without an XD tag)
__global__ static void dirs_from_flags_kernel(const uint16_t* __restrict__ flag, const int64_t* __restrict__ seq_off, int64_t r0, int64_t r1, uint8_t* __restrict__ dirs) {
    const int64_t r = r0 + blockIdx.x;
    if (r >= r1) return;
    const uint8_t d = (flag[r] & 0x10) ? DIR_R : DIR_F;
    for (int64_t k = seq_off[r] + threadIdx.x; k < seq_off[r + 1]; k += blockDim.x) dirs[k] = d;
}

// PB2 packed sequence bytes -> the store's ASCII bases + qualities; then the exceptions (bases that are not A/C/G/T, qualities the byte cannot hold)
__global__ static void unpack_seq_kernel(const uint8_t* __restrict__ seq, int64_t n, uint8_t* __restrict__ bases, uint8_t* __restrict__ quals) {
    // sixteen packed bytes per thread: one 16-byte load, two 16-byte stores (the vector path needs the three pointers 16-byte aligned; the head of a
    // push that appends at an odd offset and the tail go byte by byte)
    const uint32_t lut = 'A' | ('G' << 8) | ('C' << 16) | ((uint32_t)'T' << 24);   // allele2: AlleleType order A G C T
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i >= n) return;
    const bool aligned = ((reinterpret_cast<uintptr_t>(seq) | reinterpret_cast<uintptr_t>(bases) | reinterpret_cast<uintptr_t>(quals)) & 15) == 0;
    if (aligned && i + 16 <= n) {
        const uint4 v = *reinterpret_cast<const uint4*>(seq + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t b[4], q[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            q[k] = w[k] & 0x3f3f3f3fu;
            uint32_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) o |= ((lut >> (8 * ((w[k] >> (8 * j + 6)) & 3u))) & 0xffu) << (8 * j);
            b[k] = o;
        }
        *reinterpret_cast<uint4*>(bases + i) = make_uint4(b[0], b[1], b[2], b[3]);
        *reinterpret_cast<uint4*>(quals + i) = make_uint4(q[0], q[1], q[2], q[3]);
        return;
    }
    for (int k = 0; k < 16 && i + k < n; k++) {
        const uint32_t v = seq[i + k];
        bases[i + k] = (uint8_t)(lut >> (8 * (v >> 6)));
        quals[i + k] = (uint8_t)(v & 63u);
    }
}
__global__ static void apply_seq_exceptions_kernel(const int64_t* __restrict__ index, const uint8_t* __restrict__ eb, const uint8_t* __restrict__ eq, int64_t n_exc, int64_t lo,
                                                   int64_t hi, uint8_t* __restrict__ bases, uint8_t* __restrict__ quals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_exc) return;
    const int64_t k = index[i];
    if (k >= lo && k < hi) { bases[k] = eb[i]; quals[k] = eq[i]; }
}

// compact offsets of pb2_packed_read_batch: operations per read -> batch-relative cigar_off / seq_off, on the device
__global__ static void widen_ops_kernel(const uint8_t* __restrict__ ops, int64_t n, int64_t* __restrict__ out /* [n + 1] */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = i < n ? ops[i] : 0;
}
__global__ static void read_spans_kernel(const uint32_t* __restrict__ cigar, int64_t n_cigar, const int64_t* __restrict__ coff /* [n + 1], batch-relative */, int64_t n,
                                         int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int64_t span = 0;
    if (i < n)
        for (int64_t k = coff[i]; k < min(coff[i + 1], n_cigar); k++) {   // (counts that overrun the batch's cigar[] fail the totals check; never read past it)
            const uint32_t c = cigar[k];
            const int op = c & 15;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) span += c >> 4;   // M I S = X consume read bases
        }
    out[i] = span;
}

// A large host-to-device copy, PACED: pieces of 16 MB, at most two of them queued at the copy engine at any time (the host waits for piece i - 2
// before it issues piece i; measured with 6 concurrent jobs: 1 MB x 4 76 M, 4 MB x 3 92 M, 16 MB x 2 97 M, 32 MB x 2 97 M loci/s end to end). The engine serves what is queued in order, whatever the stream: a job that queues half a gigabyte at once makes every small
// copy of the OTHER jobs' flushes (candidates, gapped counts: a dozen per flush) wait for all of it, the jobs fall into lockstep - all copying, then all
// flushing with the link idle - and six concurrent jobs moved 40 GB/s over a link that does 55 (73 M loci/s). Paced, a small copy waits for a few pieces.
static cudaError_t h2d_in_pieces(pb2_handle* h, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    constexpr size_t kPiece = 16u << 20, kDepth = 2;
    if (bytes <= 4 * kPiece) return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess;
    size_t i = 0;
    for (size_t off = 0; off < bytes; off += kPiece, i++) {
        cudaError_t e;
        if (i >= kDepth && (e = cudaEventSynchronize(h->ev_piece[(i - kDepth) & 3])) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(static_cast<uint8_t*>(dst) + off, static_cast<const uint8_t*>(src) + off, std::min(kPiece, bytes - off), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(h->ev_piece[i & 3], st)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// seq / exc_*: the packed form of pb2_push_reads_packed (then b->bases / b->quals are NULL), else nullptr. cigar_ops: its compact-offsets form (then
// b->cigar_off / b->seq_off are NULL and ncig_total / nseq_total size the batch)
static int push_reads_impl(pb2_handle* h, const pb2_read_batch* b, const uint8_t* seq, int64_t n_exc, const int64_t* exc_index, const uint8_t* exc_base,
                           const uint8_t* exc_qual, const uint8_t* cigar_ops = nullptr, int64_t ncig_total = 0, int64_t nseq_total = 0) {
    if (!h || !b || b->n_reads < 0) return fail(h, PB2_ERR_ARG, "pb2_push_reads: bad argument");
    if (b->n_reads == 0) return PB2_OK;
    const bool compact = cigar_ops != nullptr;
    if (compact && (b->cigar_off || b->seq_off || !seq || ncig_total < 0 || nseq_total < 0)) return fail(h, PB2_ERR_ARG, "pb2_push_reads_packed: cigar_ops goes with NULL cigar_off / seq_off");
    if (!b->pos0 || !b->flag || (!compact && (!b->cigar_off || !b->seq_off)) || !b->cigar || (!seq && (!b->bases || !b->quals))) return fail(h, PB2_ERR_ARG, "pb2_push_reads: null array");
    if (seq && (n_exc < 0 || (n_exc > 0 && (!exc_index || !exc_base || !exc_qual)))) return fail(h, PB2_ERR_ARG, "pb2_push_reads_packed: bad exception list");
    CU(h, cudaSetDevice(h->device));
    nvtx_range nv("pb2_push_reads");
    Trace tr("push_reads");
    DeviceReads& R = h->reads;
    cudaStream_t st = h->stream;
    const int64_t nb = b->n_reads;
    const int64_t c_lo = compact ? 0 : b->cigar_off[0], c_hi = compact ? ncig_total : b->cigar_off[nb], s_lo = compact ? 0 : b->seq_off[0], s_hi = compact ? nseq_total : b->seq_off[nb];
    if (c_hi < c_lo || s_hi < s_lo) return fail(h, PB2_ERR_ARG, "pb2_push_reads: offsets not monotone");
    if (R.n + nb > INT32_MAX) return fail(h, PB2_ERR_ARG, "pb2_push_reads: more than 2^31 reads staged");
    const int64_t first = R.n, ncig = c_hi - c_lo, nseq = s_hi - s_lo;
    // per-base directions / collapsed summaries: optional per batch. Once any batch carried them the store holds them for every read (a stitched or
    // collapsed BAM streamed in small batches mixes tagged and untagged reads): reads pushed without get the flag-derived defaults.
    const bool want_dirs = R.has_dirs || b->base_dirs != nullptr, want_coll = R.has_collapsed || b->collapsed != nullptr;
    // amplicon names: tracked only with a threshold (Factory.ShouldTrackAmpliconCounts); reads pushed without ids have no XN tag (-1)
    const bool want_amp = h->cfg.amplicon_bias_filter >= 0 && (R.has_amplicon || b->amplicon != nullptr);
    if (want_amp && h->cfg.call_mnvs && b->amplicon != nullptr)
        return fail(h, PB2_ERR_UNSUPPORTED, "amplicon names together with call_mnvs: SupportByAmplicon of MNV-derived SNVs is not built (DESIGN.md 9)");
    CU(h, grow(h, R.pos0, (size_t)(first + nb), (size_t)first));
    CU(h, grow(h, R.end_pos, (size_t)(first + nb), (size_t)first));
    CU(h, grow(h, R.flag, (size_t)(first + nb), (size_t)first));
    CU(h, grow(h, R.cigar_off, (size_t)(first + nb + 1), (size_t)(first ? first + 1 : 0)));
    CU(h, grow(h, R.seq_off, (size_t)(first + nb + 1), (size_t)(first ? first + 1 : 0)));
    CU(h, grow(h, R.cigar, (size_t)(R.n_cigar + ncig), (size_t)R.n_cigar));
    CU(h, grow(h, R.bases, (size_t)(R.n_seq + nseq), (size_t)R.n_seq));
    CU(h, grow(h, R.quals, (size_t)(R.n_seq + nseq), (size_t)R.n_seq));
    CU(h, grow(h, R.slots, (size_t)(R.n_seq + nseq) + 32, R.n_seq ? (size_t)R.n_seq + 16 : 0));
    if (want_dirs) {
        CU(h, grow(h, R.base_dirs, (size_t)(R.n_seq + nseq), R.has_dirs ? (size_t)R.n_seq : 0));
        if (!R.has_dirs && first > 0) dirs_from_flags_kernel<<<(unsigned)first, 64, 0, st>>>(R.flag.p, R.seq_off.p, 0, first, R.base_dirs.p);
    }
    if (want_coll) {
        CU(h, grow(h, R.collapsed, (size_t)(first + nb), R.has_collapsed ? (size_t)first : 0));
        if (!R.has_collapsed && first > 0) CU(h, cudaMemsetAsync(R.collapsed.p, 0, (size_t)first, st));
    }
    if (want_amp) {
        CU(h, grow(h, R.amplicon, (size_t)(first + nb), R.has_amplicon ? (size_t)first : 0));
        if (!R.has_amplicon && first > 0) CU(h, cudaMemsetAsync(R.amplicon.p, 0xff, sizeof(int32_t) * (size_t)first, st));
    }
    if (first == 0) { CU(h, cudaMemsetAsync(R.cigar_off.p, 0, sizeof(int64_t), st)); CU(h, cudaMemsetAsync(R.seq_off.p, 0, sizeof(int64_t), st)); }
    const cudaMemcpyKind k = cudaMemcpyHostToDevice;
    CU(h, cudaMemcpyAsync(R.pos0.p + first, b->pos0, sizeof(int32_t) * (size_t)nb, k, st));
    CU(h, cudaMemcpyAsync(R.flag.p + first, b->flag, sizeof(uint16_t) * (size_t)nb, k, st));
    if (!compact) {
        CU(h, cudaMemcpyAsync(R.cigar_off.p + first + 1, b->cigar_off + 1, sizeof(int64_t) * (size_t)nb, k, st));
        CU(h, cudaMemcpyAsync(R.seq_off.p + first + 1, b->seq_off + 1, sizeof(int64_t) * (size_t)nb, k, st));
    }
    if (ncig) CU(h, cudaMemcpyAsync(R.cigar.p + R.n_cigar, b->cigar + c_lo, sizeof(uint32_t) * (size_t)ncig, k, st));
    int64_t compact_totals[2] = {ncig_total, nseq_total};
    if (compact) {   // batch-relative offsets out of the operation counts: two scans, the second over the read spans the first one's offsets delimit
        uint8_t* d_ops = nullptr;
        int64_t *d_in = nullptr, *d_coff = nullptr, *d_soff = nullptr;
        void* temp = nullptr;
        size_t tb = 0;
        CU(h, pool_alloc_t(h, &d_ops, (size_t)nb)); CU(h, pool_alloc_t(h, &d_in, (size_t)nb + 1)); CU(h, pool_alloc_t(h, &d_coff, (size_t)nb + 1)); CU(h, pool_alloc_t(h, &d_soff, (size_t)nb + 1));
        CU(h, exclusive_scan_i64(d_in, d_coff, nb + 1, nullptr, 0, &tb, st));
        CU(h, pool_alloc(h, &temp, tb + 16));
        CU(h, cudaMemcpyAsync(d_ops, cigar_ops, (size_t)nb, k, st));
        const unsigned grid = (unsigned)((nb + 1 + 255) / 256);
        widen_ops_kernel<<<grid, 256, 0, st>>>(d_ops, nb, d_in);
        CU(h, exclusive_scan_i64(d_in, d_coff, nb + 1, temp, tb, nullptr, st));
        read_spans_kernel<<<grid, 256, 0, st>>>(R.cigar.p + R.n_cigar, ncig, d_coff, nb, d_in);
        CU(h, exclusive_scan_i64(d_in, d_soff, nb + 1, temp, tb, nullptr, st));
        CU(h, cudaMemcpyAsync(R.cigar_off.p + first + 1, d_coff + 1, sizeof(int64_t) * (size_t)nb, cudaMemcpyDeviceToDevice, st));
        CU(h, cudaMemcpyAsync(R.seq_off.p + first + 1, d_soff + 1, sizeof(int64_t) * (size_t)nb, cudaMemcpyDeviceToDevice, st));
        CU(h, cudaMemcpyAsync(&compact_totals[0], d_coff + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CU(h, cudaMemcpyAsync(&compact_totals[1], d_soff + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        h->total_launches += 4;
        void* ptrs[] = {d_ops, d_in, d_coff, d_soff, temp};
        for (void* p : ptrs) pool_free(h, p);
    }
    uint8_t* d_seq = nullptr;
    int64_t* d_exc_index = nullptr;
    uint8_t *d_exc_base = nullptr, *d_exc_qual = nullptr;
    if (nseq) {
        if (seq) {
            CU(h, pool_alloc(h, (void**)&d_seq, (size_t)nseq));
            CU(h, h2d_in_pieces(h, d_seq, seq + s_lo, (size_t)nseq, st));
            unpack_seq_kernel<<<(unsigned)(((nseq + 15) / 16 + 255) / 256), 256, 0, st>>>(d_seq, nseq, R.bases.p + R.n_seq, R.quals.p + R.n_seq);
            if (n_exc > 0) {
                CU(h, pool_alloc_t(h, &d_exc_index, (size_t)n_exc)); CU(h, pool_alloc_t(h, &d_exc_base, (size_t)n_exc)); CU(h, pool_alloc_t(h, &d_exc_qual, (size_t)n_exc));
                CU(h, cudaMemcpyAsync(d_exc_index, exc_index, sizeof(int64_t) * (size_t)n_exc, k, st));
                CU(h, cudaMemcpyAsync(d_exc_base, exc_base, (size_t)n_exc, k, st));
                CU(h, cudaMemcpyAsync(d_exc_qual, exc_qual, (size_t)n_exc, k, st));
                // exception indices address the batch's seq array: rebased to the window [s_lo, s_hi) by the pointer arithmetic below
                apply_seq_exceptions_kernel<<<(unsigned)((n_exc + 255) / 256), 256, 0, st>>>(d_exc_index, d_exc_base, d_exc_qual, n_exc, s_lo, s_hi, R.bases.p + R.n_seq - s_lo,
                                                                                             R.quals.p + R.n_seq - s_lo);
            }
            h->total_launches += 2;
        } else {
            CU(h, h2d_in_pieces(h, R.bases.p + R.n_seq, b->bases + s_lo, (size_t)nseq, st));
            CU(h, h2d_in_pieces(h, R.quals.p + R.n_seq, b->quals + s_lo, (size_t)nseq, st));
        }
        if (b->base_dirs) CU(h, h2d_in_pieces(h, R.base_dirs.p + R.n_seq, b->base_dirs + s_lo, (size_t)nseq, st));
    }
    if (nseq) { CU(h, launch_reads_slots(R.bases.p + R.n_seq, R.quals.p + R.n_seq, nseq, R.slots.p + 16 + R.n_seq, st)); h->total_launches += 1; }
    if (want_coll) {
        if (b->collapsed) CU(h, cudaMemcpyAsync(R.collapsed.p + first, b->collapsed, (size_t)nb, k, st));
        else CU(h, cudaMemsetAsync(R.collapsed.p + first, 0, (size_t)nb, st));
    }
    if (want_amp) {
        if (b->amplicon) CU(h, cudaMemcpyAsync(R.amplicon.p + first, b->amplicon, sizeof(int32_t) * (size_t)nb, k, st));
        else CU(h, cudaMemsetAsync(R.amplicon.p + first, 0xff, sizeof(int32_t) * (size_t)nb, st));
    }
    // ingest: offsets rebased, Read.EndPosition, validation, the batch triggers of SmallVariantCaller.Execute
    IngestStatus* d_status = nullptr;
    int2* d_trig = nullptr;
    const int32_t trig_cap = (int32_t)std::min<int64_t>(nb, 1 << 22);
    CU(h, pool_alloc_t(h, &d_status, 1));
    CU(h, pool_alloc_t(h, &d_trig, (size_t)trig_cap));
    IngestStatus init;
    memset(&init, 0, sizeof(init));
    init.min_start = INT32_MAX;
    CU(h, cudaMemcpyAsync(d_status, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    R.n = first + nb;
    const bool had_amp = R.has_amplicon;
    R.has_dirs = want_dirs; R.has_collapsed = want_coll; R.has_amplicon = want_amp;
    ReadsView rv = R.view();
    CU(h, launch_reads_ingest(rv, (int32_t)first, R.n_cigar - c_lo, R.n_seq - s_lo, R.cigar_off.p, R.seq_off.p, R.end_pos.p, h->push_last_key, d_trig, trig_cap, d_status, h->cfg.expect_collapsed, st));
    if (want_dirs && !b->base_dirs) dirs_from_flags_kernel<<<(unsigned)nb, 64, 0, st>>>(R.flag.p, R.seq_off.p, first, first + nb, R.base_dirs.p);
    h->total_launches += 2;
    IngestStatus status;
    CU(h, cudaMemcpyAsync(&status, d_status, sizeof(status), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));   // the caller's buffers are consumed
    tr.mark("h2d+ingest");
    pool_free(h, d_seq); pool_free(h, d_exc_index); pool_free(h, d_exc_base); pool_free(h, d_exc_qual);
    auto rollback = [&]() { R.n = first; R.has_amplicon = had_amp && first > 0; pool_free(h, d_status); pool_free(h, d_trig); };
    if (compact && (compact_totals[0] != ncig_total || compact_totals[1] != nseq_total)) {
        rollback();
        return fail(h, PB2_ERR_ARG, "pb2_push_reads_packed: cigar_ops do not add up to n_cigar_total / the CIGARs' read spans to n_seq_total");
    }
    if (status.error) {
        rollback();
        switch (status.error) {
            case 1: return fail(h, PB2_ERR_ARG, "pb2_push_reads: bad CIGAR operation");
            case 2: return fail(h, PB2_ERR_ARG, "Invalid cigar: does not match length of read");   // Read.cs:603-605
            case 3: return fail(h, PB2_ERR_ARG, "Position must be greater than 0.");               // RegionStateManager.cs:363-364
            case 5: return fail(h, PB2_ERR_ARG, "The input is collapsed BAM, but a read is not a collapsed read.");   // CollapedRegionStateManager.cs:40-43
            case 6: return fail(h, PB2_ERR_UNSUPPORTED, "pb2_push_reads: an insertion or deletion longer than 65534 bases");
            default: return fail(h, PB2_ERR_ARG, "pb2_push_reads: offsets not monotone");
        }
    }
    if (status.n_triggers > trig_cap) { rollback(); return fail(h, PB2_ERR_UNSUPPORTED, "pb2_push_reads: reads are not in position order (more than 4 M block changes in one batch)"); }
    if (status.n_triggers > 0) {
        std::vector<int2> trig((size_t)status.n_triggers);
        CU(h, cudaMemcpy(trig.data(), d_trig, sizeof(int2) * trig.size(), cudaMemcpyDeviceToHost));
        std::sort(trig.begin(), trig.end(), [](const int2& a, const int2& b2) { return a.x < b2.x; });
        for (auto& t : trig) h->triggers.push_back(t.y);
    }
    pool_free(h, d_status); pool_free(h, d_trig);
    {
        const int32_t last = b->pos0[nb - 1];
        h->push_last_key = last <= 0 ? 0 : (last + 999) / 1000;
        R.last_pos0 = last;
    }
    R.n_cigar += ncig; R.n_seq += nseq;
    R.min_start = std::min(R.min_start, status.min_start);
    R.max_end = std::max(R.max_end, status.max_end);
    BatchHostView hv{seq ? nullptr : b->bases + s_lo, compact ? nullptr : b->seq_off, s_lo};
    const int rc = explicit_find_candidates(h, (size_t)first, seq ? nullptr : &hv);
    tr.mark("candidates");
    return rc;
}

extern "C" int pb2_push_reads(pb2_handle* h, const pb2_read_batch* b) { return push_reads_impl(h, b, nullptr, 0, nullptr, nullptr, nullptr); }
extern "C" int pb2_push_reads_packed(pb2_handle* h, const pb2_packed_read_batch* p) {
    if (!h || !p) return fail(h, PB2_ERR_ARG, "pb2_push_reads_packed: bad argument");
    pb2_read_batch b;
    memset(&b, 0, sizeof(b));
    b.n_reads = p->n_reads; b.pos0 = p->pos0; b.flag = p->flag; b.cigar_off = p->cigar_off; b.cigar = p->cigar; b.seq_off = p->seq_off;
    b.base_dirs = p->base_dirs; b.collapsed = p->collapsed; b.amplicon = p->amplicon;
    if (p->n_reads > 0 && !p->seq) return fail(h, PB2_ERR_ARG, "pb2_push_reads_packed: null array");
    if (p->cigar_ops != nullptr) return push_reads_impl(h, &b, p->seq, p->n_exceptions, p->exc_index, p->exc_base, p->exc_qual, p->cigar_ops, p->n_cigar_total, p->n_seq_total);
    return push_reads_impl(h, &b, p->seq, p->n_exceptions, p->exc_index, p->exc_base, p->exc_qual);
}
// Host helper: bases + qualities -> the packed bytes of pb2_push_reads_packed and their exception list. Returns the number of exceptions (which may
// exceed exc_capacity: then only the first exc_capacity were stored and the caller retries with larger arrays), or < 0 on a bad argument.
extern "C" int64_t pb2_pack_reads(const uint8_t* bases, const uint8_t* quals, int64_t n, uint8_t* seq, int64_t* exc_index, uint8_t* exc_base, uint8_t* exc_qual,
                                  int64_t exc_capacity) {
    if (n < 0 || (n > 0 && (!bases || !quals || !seq))) return PB2_ERR_ARG;
    int64_t n_exc = 0;
    for (int64_t i = 0; i < n; i++) {
        const uint8_t b = bases[i], q = quals[i];
        const int a2 = b == 'A' ? 0 : b == 'G' ? 1 : b == 'C' ? 2 : b == 'T' ? 3 : -1;
        const uint8_t v = (a2 < 0 || q > 63) ? 0 : (uint8_t)((a2 << 6) | q);
        seq[i] = v;
        if (v == 0) {
            if (n_exc < exc_capacity && exc_index && exc_base && exc_qual) { exc_index[n_exc] = i; exc_base[n_exc] = b; exc_qual[n_exc] = q; }
            n_exc++;
        }
    }
    return n_exc;
}

// Host helper of the locus-major path: the three planes of pb2_pileup_csr -> PB2_LAYOUT_PACKED2 (two bytes per entry: the anchor bin rides in the spare bits of
// the code and quality bytes) plus the sparse list of entries that carry candidate flags (code bits 5-7). With offsets + ref_bases, flags on entries whose
// base is the reference base of their locus are left out: such a base raises no SNV candidate (CandidateVariantFinder.cs:112-141), its open-end flags say
// nothing (staging clears them as well). Returns the number of flagged entries (which may exceed flag_capacity: then only the first flag_capacity were
// stored), PB2_ERR_ARG for a bad argument, PB2_ERR_UNSUPPORTED when an anchor byte carries a collapsed-read type (bits 4-7: the packed form has no room).
extern "C" int64_t pb2_pack_pileup(const uint8_t* code, const uint8_t* qual, const uint8_t* anchor, int64_t n_entries, const int64_t* offsets, const uint8_t* ref_bases,
                                   int64_t n_loci, uint8_t* pcode, uint8_t* pqual, int64_t* flag_index, uint8_t* flag_bits, int64_t flag_capacity) {
    if (n_entries < 0 || (n_entries > 0 && (!code || !qual || !anchor || !pcode || !pqual)) || flag_capacity < 0) return PB2_ERR_ARG;
    const bool by_locus = offsets != nullptr && ref_bases != nullptr && n_loci > 0;
    if (by_locus && offsets[n_loci] != n_entries) return PB2_ERR_ARG;
    int64_t n_flags = 0, locus = 0;
    for (int64_t i = 0; i < n_entries; i++) {
        const uint8_t c = code[i], q = qual[i], a = anchor[i];
        if (a >> 4) return PB2_ERR_UNSUPPORTED;
        pcode[i] = (uint8_t)((c & 0x1f) | ((a & 7) << 5));
        pqual[i] = (uint8_t)((q & 0x7f) | ((a >> 3) << 7));
        if (c & 0xe0) {
            bool keep = true;
            if (by_locus) {
                while (locus + 1 <= n_loci && offsets[locus + 1] <= i) locus++;
                const uint8_t rb = ref_bases[locus];
                const int ref_allele = rb == 'A' ? 0 : rb == 'G' ? 1 : rb == 'C' ? 2 : rb == 'T' ? 3 : 4;
                keep = (c & 7) != ref_allele;
            }
            if (keep) {
                if (n_flags < flag_capacity && flag_index && flag_bits) { flag_index[n_flags] = i; flag_bits[n_flags] = (uint8_t)(c & 0xe0); }
                n_flags++;
            }
        }
    }
    return n_flags;
}

// ------------------------------------------------------------------------------------------------ interval sharding
// BaseGenomeProcessor shards by chromosome and concatenates in genome order (BaseGenomeProcessor.cs:60-72, GenomeProcessor.cs:156-186); within a chromosome the
// loci are independent once the counts are complete, so a chromosome is cut into interval shards at 1000-bp block boundaries (the batches of
// RegionStateManager are block-aligned). A shard stages a HALO on either side - two blocks plus the longest read span: what reaches across a cut is an
// allele's far end point (CoverageCalculator.cs:19-47), MNV leftovers that move into the next block (AlleleCaller.cs:91-92), the collapsable candidates a
// batch pulls from the following block (RegionStateManager.cs:441-457) and the gapped-MNV reference take-away (AlleleCaller.cs:94), all within one read
// span of the cut - and emits only the positions it owns.
extern "C" int pb2_shard_plan(const int32_t* pos0, int64_t n_reads, int32_t first_position, int32_t last_position, int32_t max_read_span, int32_t n_shards, pb2_shard* out) {
    if (n_shards < 1 || !out || last_position < first_position || first_position < 1 || max_read_span < 0 || n_reads < 0 || (n_reads > 0 && !pos0)) return PB2_ERR_ARG;
    const int32_t halo = 2000 + (max_read_span + 999) / 1000 * 1000;
    std::vector<int32_t> cut((size_t)n_shards + 1);   // shard i owns (cut[i], cut[i + 1]]
    cut[0] = first_position - 1;
    cut[(size_t)n_shards] = last_position;
    for (int32_t i = 1; i < n_shards; i++) {
        int64_t p;
        if (n_reads > 0) p = (int64_t)pos0[(size_t)(n_reads * i / n_shards)] + 1;                                      // balanced by reads (pos0 is sorted)
        else p = first_position - 1 + ((int64_t)last_position - first_position + 1) * i / n_shards;                      // balanced by positions
        p = (p + 500) / 1000 * 1000;                                                                                     // to a block boundary
        p = std::max<int64_t>(p, cut[(size_t)i - 1]);
        cut[(size_t)i] = (int32_t)std::min<int64_t>(p, last_position);
    }
    for (int32_t i = 0; i < n_shards; i++) {
        pb2_shard& s = out[i];
        s.own_lo = cut[(size_t)i] + 1; s.own_hi = cut[(size_t)i + 1];
        s.stage_lo = std::max(1, s.own_lo - halo);
        s.stage_hi = (int32_t)std::min<int64_t>((int64_t)s.own_hi + halo, INT32_MAX);
        s.read_first = 0; s.read_end = n_reads;
        if (n_reads > 0) {   // the position-sorted reads that can touch [stage_lo, stage_hi]
            s.read_first = std::lower_bound(pos0, pos0 + n_reads, s.stage_lo - 1 - max_read_span) - pos0;
            s.read_end = std::upper_bound(pos0, pos0 + n_reads, s.stage_hi - 1) - pos0;
        }
        if (s.own_hi < s.own_lo) { s.read_first = s.read_end = 0; }
    }
    return PB2_OK;
}
extern "C" int pb2_set_owned_range(pb2_handle* h, int32_t own_lo, int32_t own_hi) {
    if (!h || (own_hi != 0 && (own_lo < 1 || own_hi < own_lo))) return fail(h, PB2_ERR_ARG, "pb2_set_owned_range: bad range");
    h->own_lo = own_hi ? own_lo : 0; h->own_hi = own_hi;
    derive_config(h);
    explicit_release_resident(h);
    release_resident_graph(h);
    return PB2_OK;
}

extern "C" int pb2_totals(pb2_handle* h, int64_t* total_collapsed) {
    if (!h) return PB2_ERR_ARG;
    if (total_collapsed) *total_collapsed = h->total_collapsed;
    return PB2_OK;
}

template <class T>
static cudaError_t upload(T** d, const std::vector<T>& v, cudaStream_t st) {
    cudaError_t e = cudaMalloc(d, std::max<size_t>(v.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (!v.empty()) e = cudaMemcpyAsync(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    return e;
}

// the outputs every segment needs: record streams, queues, counters (the flagged-entry side list is allocated by the staging itself)
static int alloc_segment_outputs(pb2_handle* h, Segment& s) {
    if (h->cfg.output_gvcf) {
        s.alloc_ref = sizeof(pb2_call_record) * (size_t)s.n_loci;
        CU(h, pool_alloc(h, (void**)&s.ref_records, s.alloc_ref));
        CU(h, pool_alloc_t(h, &s.ref_valid, (size_t)s.n_loci));
    }
    s.var_capacity = std::max<int64_t>(1024, s.n_loci);
    s.alloc_var = sizeof(pb2_call_record) * (size_t)s.var_capacity;
    CU(h, pool_alloc(h, (void**)&s.var_records, s.alloc_var));
    s.pending_capacity = std::max<int64_t>(1024, s.n_loci);
    s.alloc_pending = sizeof(PendingLocus) * (size_t)s.pending_capacity;
    CU(h, pool_alloc(h, (void**)&s.pending, s.alloc_pending));
    return PB2_OK;
}

// can the reads of this handle be staged in the PVERT form (pb2_pvert.cuh)?
static bool pvert_eligible(const pb2_handle* h) {
    return h->cfg.min_base_call_quality >= 2 && h->cfg.min_base_call_quality <= 63 && !h->cfg.want_sum_base_quality && h->cfg.noise_model != 1 && h->cfg.reserved[1] != 9;
}

PvertPileup pvert_view(const Segment& s) {
    PvertPileup pv;
    memset(&pv, 0, sizeof(pv));
    pv.data = s.pv_data; pv.row_meta = s.pv_row_meta; pv.tile_row0 = s.pv_tile_row0; pv.cls_end = s.pv_cls_end; pv.n_classes = s.pv_classes;
    pv.ref_base = s.ref_base; pv.positions = s.positions; pv.first_position = s.first_position; pv.n_loci = s.n_loci; pv.n_tiles = s.n_tiles;
    return pv;
}

// Build a segment from the staged reads for reference positions cleared_from .. cleared_end (INT32_MAX = everything). temporary: built inside
// pb2_flush / pb2_get_counts and freed when they return; otherwise it stays staged (pb2_stage_reads -> pb2_call_resident).
static int stage_reads_segment(pb2_handle* h, int32_t cleared_end, int32_t cleared_from, bool temporary = true) {
    DeviceReads& R = h->reads;
    if (R.size() == 0) return PB2_OK;
    nvtx_range nv("stage_reads");
    int32_t lo = R.min_start, hi = R.max_end;
    if (h->have_intervals) {   // interval positions are reported for every position of a block some read touched, covered or not
        lo = ((lo - 1) / 1000) * 1000 + 1;
        hi = (int32_t)std::min<int64_t>(((int64_t)(hi - 1) / 1000 + 1) * 1000, INT32_MAX);
    }
    lo = std::max(lo, cleared_from);
    hi = std::min(hi, cleared_end);
    if (hi < lo) return PB2_OK;
    cudaStream_t st = h->stream;
    const int64_t span = (int64_t)hi - lo + 1;
    // loci: the whole span, or — with intervals — the interval positions inside 1000-bp blocks some read touched
    // (reference candidates are only generated for existing blocks: RegionStateManager.cs:295-314, RegionState.cs:393-399)
    std::vector<int32_t> positions, index_of_pos;
    // without intervals the reads of one flush can still lie far apart (an amplicon panel run without an interval file): past a million positions of
    // span the touched blocks are looked at, and if they are the smaller half only their positions are staged (the reference only ever creates the
    // blocks reads touch, RegionStateManager.cs:361-383) - through the same positions[] / index_of_pos indirection as interval runs
    bool listed = h->have_intervals;
    if (h->have_intervals || span > (1 << 20)) {
        const int b0 = (lo - 1) / 1000, nbk = (hi - 1) / 1000 - b0 + 1;
        std::vector<uint32_t> bits((size_t)(nbk + 31) / 32, 0);
        uint32_t* d_bits = nullptr;
        CU(h, pool_alloc_t(h, &d_bits, bits.size()));
        CU(h, cudaMemsetAsync(d_bits, 0, sizeof(uint32_t) * bits.size(), st));
        // block key of position p is (p + 999) / 1000 = (p - 1) / 1000 + 1
        CU(h, launch_reads_block_bitmap(R.pos0.p, R.end_pos.p, R.n, lo - 1, b0 + 1, nbk, d_bits, st));
        CU(h, cudaMemcpyAsync(bits.data(), d_bits, sizeof(uint32_t) * bits.size(), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
        pool_free(h, d_bits);
        auto touched = [&](int64_t p) { const int k = (int)((p - 1) / 1000) - b0; return k >= 0 && k < nbk && ((bits[(size_t)k >> 5] >> (k & 31)) & 1u); };
        if (!h->have_intervals) {
            int64_t n_touched = 0;
            for (uint32_t w : bits) n_touched += __builtin_popcount(w);
            listed = n_touched * 1000 * 2 < span;
        }
        if (listed) {
            index_of_pos.assign((size_t)span, -1);
            std::vector<std::pair<int32_t, int32_t>> iv;
            if (h->have_intervals) for (size_t i = 0; i < h->iv_start.size(); i++) iv.push_back({h->iv_start[i], h->iv_end[i]});
            else iv.push_back({lo, hi});
            std::sort(iv.begin(), iv.end());
            for (auto& v : iv)
                for (int64_t p = std::max<int64_t>(v.first, lo); p <= std::min<int64_t>(v.second, hi); p++)
                    if (index_of_pos[(size_t)(p - lo)] < 0 && touched(p)) index_of_pos[(size_t)(p - lo)] = 0;
            for (int64_t k = 0; k < span; k++)
                if (index_of_pos[(size_t)k] == 0) { index_of_pos[(size_t)k] = (int32_t)positions.size(); positions.push_back((int32_t)(lo + k)); }
            if (positions.empty()) return PB2_OK;
        }
    }
    const int64_t n_loci = listed ? (int64_t)positions.size() : span;
    int32_t *d_index = nullptr, *d_index_ge = nullptr;
    Segment s;
    s.n_loci = n_loci;
    s.n_tiles = (int32_t)((n_loci + kTileLoci - 1) / kTileLoci);
    s.first_position = lo;
    s.has_positions = listed;
    s.temporary = temporary;
    if (listed) {
        CU(h, pool_alloc_t(h, &d_index, index_of_pos.size()));
        CU(h, cudaMemcpyAsync(d_index, index_of_pos.data(), sizeof(int32_t) * index_of_pos.size(), cudaMemcpyHostToDevice, st));
        std::vector<int32_t> index_ge((size_t)span + 1);
        index_ge[(size_t)span] = (int32_t)positions.size();
        for (int64_t k = span - 1; k >= 0; k--) index_ge[(size_t)k] = index_of_pos[(size_t)k] >= 0 ? index_of_pos[(size_t)k] : index_ge[(size_t)k + 1];
        CU(h, pool_alloc_t(h, &d_index_ge, index_ge.size()));
        CU(h, cudaMemcpyAsync(d_index_ge, index_ge.data(), sizeof(int32_t) * index_ge.size(), cudaMemcpyHostToDevice, st));
        CU(h, cudaStreamSynchronize(st));   // index_ge is a local
        CU(h, pool_alloc_t(h, &s.positions, positions.size()));
        CU(h, cudaMemcpyAsync(s.positions, positions.data(), sizeof(int32_t) * positions.size(), cudaMemcpyHostToDevice, st));
        s.h_positions = positions;
    }
    ReadsView rv = R.view();
    RegionView rg{lo, hi, d_index, h->d_chr, h->chr_len, h->dcfg.min_bq, h->cfg.expect_collapsed, d_index_ge, s.positions, n_loci};
    if (!pvert_eligible(h)) {
        if (rv.amplicon != nullptr) {
            pool_free(h, d_index); pool_free(h, d_index_ge); pool_free(h, s.positions);
            return fail(h, PB2_ERR_UNSUPPORTED, "amplicon names need the PVERT pileup: minimum base-call quality in [2, 63], flat noise model, no quality sums");
        }
        // quality sums / unusual quality bars: the PTILE32 form, through the locus-major entry list (reads_count / reads_emit -> push_common)
        unsigned int *d_depth = nullptr, *d_cursor = nullptr;
        int64_t *d_off = nullptr, *d_depth64 = nullptr;
        uint8_t *d_code = nullptr, *d_qual = nullptr, *d_anch = nullptr, *d_ref = nullptr;
        void* temp = nullptr;
        auto cleanup = [&]() {
            void* ptrs[] = {d_depth, d_cursor, d_off, d_depth64, d_code, d_qual, d_anch, d_ref, temp, d_index, d_index_ge, s.positions};
            for (void* p : ptrs) pool_free(h, p);
        };
#define CUC(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { h->error = std::string(#expr) + ": " + cudaGetErrorString(_e); cleanup(); return PB2_ERR_CUDA; } } while (0)
        CUC(pool_alloc_t(h, &d_depth, (size_t)n_loci));
        CUC(pool_alloc_t(h, &d_cursor, (size_t)n_loci));
        CUC(cudaMemsetAsync(d_depth, 0, sizeof(unsigned int) * (size_t)n_loci, st));
        CUC(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned int) * (size_t)n_loci, st));
        CUC(launch_reads_count(rv, rg, d_depth, st));
        CUC(pool_alloc_t(h, &d_off, (size_t)(n_loci + 1)));
        CUC(pool_alloc_t(h, &d_depth64, (size_t)(n_loci + 1)));
        CUC(launch_depth_to_i64(d_depth, d_depth64, n_loci, st));
        size_t tb = 0;
        CUC(exclusive_scan_i64(d_depth64, d_off, n_loci + 1, nullptr, 0, &tb, st));
        CUC(pool_alloc(h, &temp, std::max<size_t>(tb, 16)));
        CUC(exclusive_scan_i64(d_depth64, d_off, n_loci + 1, temp, tb, nullptr, st));
        int64_t n_entries = 0;
        CUC(cudaMemcpyAsync(&n_entries, d_off + n_loci, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CUC(cudaStreamSynchronize(st));
        CUC(pool_alloc(h, (void**)&d_code, (size_t)std::max<int64_t>(n_entries, 1)));
        CUC(pool_alloc(h, (void**)&d_qual, (size_t)std::max<int64_t>(n_entries, 1)));
        CUC(pool_alloc(h, (void**)&d_anch, (size_t)std::max<int64_t>(n_entries, 1)));
        CUC(launch_reads_emit(rv, rg, d_off, d_cursor, d_code, d_qual, d_anch, st));
        CUC(pool_alloc(h, (void**)&d_ref, (size_t)n_loci));
        CUC(launch_pvert_ref_bases(h->d_chr, h->chr_len, s.positions, lo, n_loci, d_ref, nullptr, st));
        h->total_launches += 6;
        pb2_pileup_csr csr;
        memset(&csr, 0, sizeof(csr));
        csr.n_loci = n_loci; csr.first_position = lo; csr.positions = s.positions; csr.offsets = d_off;
        csr.code = d_code; csr.qual = d_qual; csr.anchor = d_anch; csr.ref_bases = d_ref;
        const int rc = push_common(h, &csr, true);
        cleanup();
#undef CUC
        if (rc != PB2_OK) return rc;
        h->segs.back().temporary = temporary;
        return PB2_OK;
    }
    // ---- PVERT: rows per (tile, class), layout, fill, bit transposition
    CU(h, cudaEventRecord(h->ev_stage0, st));
    const int nc = h->cfg.expect_collapsed ? kPvClassesCollapsed : kPvClassesPlain;
    s.pv_classes = nc;
    int32_t* d_cursor = nullptr;
    int64_t* tile_rows = nullptr;
    void* temp = nullptr;
    const size_t n_cls = (size_t)s.n_tiles * (size_t)nc;
    CU(h, pool_alloc_t(h, &s.pv_cls_end, n_cls));
    CU(h, pool_alloc_t(h, &d_cursor, n_cls));
    CU(h, pool_alloc_t(h, &tile_rows, (size_t)s.n_tiles + 1));
    CU(h, pool_alloc_t(h, &s.pv_tile_row0, (size_t)s.n_tiles + 1));
    CU(h, cudaMemsetAsync(s.pv_cls_end, 0, sizeof(int32_t) * n_cls, st));
    CU(h, cudaMemsetAsync(d_cursor, 0, sizeof(int32_t) * n_cls, st));
    int32_t* d_complex = nullptr;
    CU(h, pool_alloc_t(h, &d_complex, (size_t)R.n + 1));
    CU(h, cudaMemsetAsync(d_complex, 0, sizeof(int32_t), st));
    CU(h, launch_pvert_count(rv, rg, R.end_pos.p, nc, s.pv_cls_end, d_complex, st));
    CU(h, launch_pvert_layout(s.pv_cls_end, s.n_tiles, nc, tile_rows, st));
    size_t tb = 0;
    CU(h, exclusive_scan_i64(tile_rows, s.pv_tile_row0, s.n_tiles + 1, nullptr, 0, &tb, st));
    CU(h, pool_alloc(h, &temp, std::max<size_t>(tb, 16)));
    CU(h, exclusive_scan_i64(tile_rows, s.pv_tile_row0, s.n_tiles + 1, temp, tb, nullptr, st));
    int32_t n_complex = 0;
    CU(h, cudaMemcpyAsync(&s.pv_rows, s.pv_tile_row0 + s.n_tiles, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(&n_complex, d_complex, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    const size_t data_bytes = (size_t)std::max<int64_t>(s.pv_rows, 32) * 32;
    CU(h, pool_alloc(h, (void**)&s.pv_data, data_bytes + 4096));
    CU(h, pool_alloc_t(h, &s.pv_row_meta, (size_t)std::max<int64_t>(s.pv_rows, 32)));
    if (rv.amplicon != nullptr) {   // amplicon tracking: the name id of every row's read (padding rows: none)
        CU(h, pool_alloc_t(h, &s.pv_row_amp, (size_t)std::max<int64_t>(s.pv_rows, 32)));
        CU(h, cudaMemsetAsync(s.pv_row_amp, 0xff, sizeof(int32_t) * (size_t)std::max<int64_t>(s.pv_rows, 32), st));
    }
    CU(h, cudaMemsetAsync(s.pv_data, 0, data_bytes + 4096, st));
    s.exc_capacity = std::max<int64_t>(1 << 20, R.n_seq / 128);
    CU(h, pool_alloc_t(h, &s.exc_entries, 2 * (size_t)s.exc_capacity));
    CU(h, pool_alloc_t(h, &s.counters, 5));   // variants, -, pending, flagged entries, amplicon status
    CU(h, cudaMemsetAsync(s.counters, 0, sizeof(unsigned long long) * 5, st));
    uint8_t* d_ref_slot = nullptr;
    CU(h, pool_alloc_t(h, &s.ref_base, (size_t)n_loci));
    CU(h, pool_alloc_t(h, &d_ref_slot, (size_t)s.n_tiles * 32));
    CU(h, launch_pvert_ref_bases(h->d_chr, h->chr_len, s.positions, lo, n_loci, s.ref_base, d_ref_slot, st));
    CU(h, launch_pvert_fill(rv, rg, R.end_pos.p, nc, s.pv_tile_row0, s.pv_cls_end, d_cursor, s.pv_data, s.pv_row_meta, s.pv_row_amp, s.exc_entries, s.counters + 3, s.exc_capacity, d_ref_slot,
                            d_complex, n_complex, st));
    CU(h, launch_pvert_transpose(s.pv_data, s.pv_rows / 32, st));
    pool_free(h, d_ref_slot);
    CU(h, cudaEventRecord(h->ev_stage1, st));
    h->total_launches += 6;
    s.n_entries = R.n_seq;
    h->last_stage_bytes = (int64_t)s.pv_rows * 32 + (int64_t)n_cls * 4 + ((int64_t)s.n_tiles + 1) * 8 + n_loci;
    h->last_stage_rows = s.pv_rows;
    h->have_stage_events = true;
    const int rc = alloc_segment_outputs(h, s);
    pool_free(h, d_cursor); pool_free(h, tile_rows); pool_free(h, temp); pool_free(h, d_index); pool_free(h, d_index_ge); pool_free(h, d_complex);
    if (rc != PB2_OK) return rc;
    h->segs.push_back(std::move(s));
    return PB2_OK;
}

// IStateManager view for hosts that keep the whole chromosome's reads on the device: stages everything pushed so far as one resident segment, so
// that pb2_call_resident / pb2_resident_results can run on it (the reads stay staged; pb2_flush still works and re-stages what it needs).
extern "C" int pb2_stage_reads(pb2_handle* h) {
    if (!h) return PB2_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    release_resident_graph(h);
    explicit_release_resident(h);
    for (size_t i = 0; i < h->segs.size();) {
        if (h->segs[i].from_reads) { free_segment(h, h->segs[i]); h->segs.erase(h->segs.begin() + (long)i); } else i++;
    }
    const size_t before = h->segs.size();
    const int rc = stage_reads_segment(h, INT32_MAX, h->cleared_through + 1, false);
    if (rc != PB2_OK) return rc;
    for (size_t i = before; i < h->segs.size(); i++) h->segs[i].from_reads = true;
    CU(h, cudaStreamSynchronize(h->stream));
    return PB2_OK;
}

// One resident step over one segment: counters reset; the explicit pass (two small latency-bound kernels) on the side stream ahead of the hot kernel so
// that it runs next to it; join; prune; counters to pinned memory. The sequence is captured once into a CUDA graph and replayed (one launch call per
// step instead of a dozen API calls); if capture is not possible the same sequence is issued directly.
static int resident_step_enqueue(pb2_handle* h, Segment& s, bool with_explicit, bool capturing) {
    cudaStream_t st = h->stream;
    CU(h, cudaMemsetAsync(s.counters, 0, sizeof(unsigned long long) * 3, st));
    if (with_explicit) {
        CU(h, cudaEventRecord(h->ev_scattered[0], st));
        CU(h, cudaStreamWaitEvent(h->copy_stream, h->ev_scattered[0], 0));
        const int rc = explicit_call_resident(h, s, h->copy_stream);
        if (rc != PB2_OK) return rc;
        CU(h, cudaEventRecord(h->ev_copied[0], h->copy_stream));
    }
    const int rc = enqueue_segment(h, s, nullptr, nullptr, nullptr, false, capturing);
    if (rc != PB2_OK) return rc;
    if (with_explicit) {
        CU(h, cudaStreamWaitEvent(st, h->ev_copied[0], 0));
        const int rc2 = explicit_prune_resident(h, s);
        if (rc2 != PB2_OK) return rc2;
    }
    { const int rc3 = amplicon_pass(h, s, st); if (rc3 != PB2_OK) return rc3; }
    CU(h, cudaMemcpyAsync(h->h_counters, s.counters, sizeof(unsigned long long) * 5, cudaMemcpyDeviceToHost, st));
    return PB2_OK;
}
// ---- the job's record sink (multi-GPU gather): [n_slots int64 counts | n_slots x slot_records records]; a step's variant stream lands in slot step % n_slots
__global__ static void sink_copy_kernel(const pb2_call_record* __restrict__ var, const unsigned long long* __restrict__ counters, long long* __restrict__ count_out,
                                        pb2_call_record* __restrict__ slot, int64_t slot_records) {
    const int64_t n = min((int64_t)counters[0], slot_records);
    if (blockIdx.x == 0 && threadIdx.x == 0) *count_out = (long long)n;
    // 8-byte words: the record blocks start 8 * n_slots bytes into the caller's buffer, which is 16-byte aligned only for an even number of slots
    const uint2* src = reinterpret_cast<const uint2*>(var);
    uint2* dst = reinterpret_cast<uint2*>(slot);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * 12; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ static void sink_keys_kernel(const long long* __restrict__ counts, const pb2_call_record* __restrict__ slots, int64_t slot_records, int32_t n_slots,
                                        unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slot_records * n_slots) return;
    const int64_t slot = i / slot_records, k = i % slot_records;
    const bool valid = k < counts[slot];
    keys[i] = ((unsigned long long)slot << 33) | (valid ? (unsigned long long)(uint32_t)slots[i].position : 0x1ffffffffull);
    idx[i] = (uint32_t)i;
}
__global__ static void sink_gather_kernel(const pb2_call_record* __restrict__ in, const uint32_t* __restrict__ idx, int64_t n, pb2_call_record* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 12) return;
    reinterpret_cast<uint2*>(out)[t] = reinterpret_cast<const uint2*>(in)[(int64_t)idx[t / 12] * 12 + t % 12];
}
cudaError_t sink_sort_pairs(void* temp, size_t& temp_bytes, const unsigned long long* kin, unsigned long long* kout, const uint32_t* vin, uint32_t* vout, int64_t n, cudaStream_t st);

extern "C" int pb2_set_resident_sink(pb2_handle* h, void* device_buffer, int64_t slot_records, int32_t n_slots) {
    if (!h || (device_buffer && (slot_records < 1 || n_slots < 1 || (reinterpret_cast<uintptr_t>(device_buffer) & 7))))
        return fail(h, PB2_ERR_ARG, "pb2_set_resident_sink: bad argument (the buffer must be 8-byte aligned)");
    h->sink = device_buffer; h->sink_slot_records = slot_records; h->sink_slots = device_buffer ? n_slots : 0; h->sink_next = 0;
    return PB2_OK;
}
static int sink_append(pb2_handle* h, Segment& s) {
    if (!h->sink) return PB2_OK;
    const int32_t slot = (int32_t)(h->sink_next++ % h->sink_slots);
    long long* counts = reinterpret_cast<long long*>(h->sink);
    pb2_call_record* slots = reinterpret_cast<pb2_call_record*>(reinterpret_cast<uint8_t*>(h->sink) + 8 * (size_t)h->sink_slots);
    const int blocks = (int)std::min<int64_t>((h->sink_slot_records * 12 + 255) / 256, 2048);
    sink_copy_kernel<<<blocks, 256, 0, h->stream>>>(s.var_records, s.counters, counts + slot, slots + (int64_t)slot * h->sink_slot_records, h->sink_slot_records);
    h->total_launches += 1;
    return cudaGetLastError() == cudaSuccess ? PB2_OK : fail(h, PB2_ERR_CUDA, "sink_copy_kernel launch failed");
}
// Orders the records of every slot by position on the device (AlleleCaller orders its output by position: AlleleCaller.cs:96-140; alleles of one position
// keep their emission order, as the reference's OrderBy is applied per position by the host that writes them)
extern "C" int pb2_sink_sort(pb2_handle* h) {
    if (!h || !h->sink) return fail(h, PB2_ERR_STATE, "pb2_sink_sort: no sink set");
    CU(h, cudaSetDevice(h->device));
    const int64_t n = h->sink_slot_records * h->sink_slots;
    long long* counts = reinterpret_cast<long long*>(h->sink);
    pb2_call_record* slots = reinterpret_cast<pb2_call_record*>(reinterpret_cast<uint8_t*>(h->sink) + 8 * (size_t)h->sink_slots);
    unsigned long long *k0 = nullptr, *k1 = nullptr;
    uint32_t *i0 = nullptr, *i1 = nullptr;
    pb2_call_record* tmp = nullptr;
    void* temp = nullptr;
    size_t tb = 0;
    CU(h, sink_sort_pairs(nullptr, tb, k0, k1, i0, i1, n, h->stream));
    CU(h, pool_alloc_t(h, &k0, (size_t)n)); CU(h, pool_alloc_t(h, &k1, (size_t)n)); CU(h, pool_alloc_t(h, &i0, (size_t)n)); CU(h, pool_alloc_t(h, &i1, (size_t)n));
    CU(h, pool_alloc_t(h, &tmp, (size_t)n)); CU(h, pool_alloc(h, &temp, tb + 16));
    sink_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(counts, slots, h->sink_slot_records, h->sink_slots, k0, i0);
    CU(h, sink_sort_pairs(temp, tb, k0, k1, i0, i1, n, h->stream));
    sink_gather_kernel<<<(unsigned)((n * 12 + 255) / 256), 256, 0, h->stream>>>(slots, i1, n, tmp);
    CU(h, cudaMemcpyAsync(slots, tmp, sizeof(pb2_call_record) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
    h->total_launches += 4;
    void* ptrs[] = {k0, k1, i0, i1, tmp, temp};
    for (void* p : ptrs) pool_free(h, p);
    return PB2_OK;
}

static int resident_step(pb2_handle* h, Segment& s, bool with_explicit, bool sync = true) {
    cudaStream_t st = h->stream;
    if (h->resident_graph == nullptr && !h->resident_graph_failed) {
        const int64_t launches_before = h->total_launches;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            ok = resident_step_enqueue(h, s, with_explicit, true) == PB2_OK;
            ok = (cudaStreamEndCapture(st, &graph) == cudaSuccess) && ok && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&h->resident_graph, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        h->resident_graph_launches = h->total_launches - launches_before + 2;   // + hot kernel and overflow scorer (counted by finish_segment otherwise)
        h->total_launches = launches_before;
        if (!ok) { cudaGetLastError(); h->resident_graph = nullptr; h->resident_graph_failed = true; }
    }
    if (h->resident_graph != nullptr) {
        CU(h, cudaGraphLaunch(h->resident_graph, st));
        h->total_launches += h->resident_graph_launches - 2;
    } else {
        const int rc = resident_step_enqueue(h, s, with_explicit, false);
        if (rc != PB2_OK) return rc;
    }
    { const int rc = sink_append(h, s); if (rc != PB2_OK) return rc; }
    if (!sync) { h->hot_launches += 1; h->total_launches += 2; return PB2_OK; }
    return finish_segment(h, s, true);
}
// One resident step without a host synchronisation: the step (a CUDA graph) and the copy of its variant records into the next sink slot are enqueued and
// the call returns. Needs one synchronous pb2_call_resident before (it builds the plan and the graph) and a sink (pb2_set_resident_sink).
extern "C" int pb2_call_resident_async(pb2_handle* h) {
    if (!h) return PB2_ERR_ARG;
    if (h->segs.size() != 1 || h->resident_graph == nullptr || (!h->cands.empty() && !explicit_resident_ready(h)))
        return fail(h, PB2_ERR_STATE, "pb2_call_resident_async: call pb2_call_resident once first (one staged segment)");
    CU(h, cudaSetDevice(h->device));
    return resident_step(h, h->segs[0], !h->cands.empty(), false);
}
extern "C" int pb2_resident_sync(pb2_handle* h, int64_t* n_records_last) {
    if (!h) return PB2_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    if (!h->segs.empty()) {
        Segment& s = h->segs[0];
        s.h_var_count = h->h_counters[0]; s.h_exc_count = h->h_counters[3]; s.called = true;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->hot_ms += ms * 1;   // (the last step's kernel time; the steps before are counted as launches)
        if (n_records_last) *n_records_last = (int64_t)s.h_var_count;
        if ((int64_t)s.h_var_count > s.var_capacity) return fail(h, PB2_ERR_NOMEM, "variant record buffer overflow");
        if (s.pv_row_amp != nullptr && h->h_counters[4] != 0) return fail(h, PB2_ERR_ARG, kTooManyAmplicons);
    }
    return PB2_OK;
}

extern "C" int pb2_call_resident(pb2_handle* h, int64_t* n_records) {
    if (!h) return PB2_ERR_ARG;
    if (h->dcfg.ploidy != PLOIDY_SOMATIC || h->cfg.ploidy == PLOIDY_DIPLOID)
        return fail(h, PB2_ERR_UNSUPPORTED, "pb2_call_resident: the germline genotypers assign one genotype per locus when pb2_flush merges the alleles; use pb2_flush");
    CU(h, cudaSetDevice(h->device));
    int64_t total = 0;
    if (!h->cands.empty()) {
        // explicit candidates: gathered, scored and appended to the segment's variant stream on the device, behind the hot kernel on the same
        // stream; one synchronize for the whole step
        if (h->segs.size() != 1) return fail(h, PB2_ERR_UNSUPPORTED, "pb2_call_resident with explicit candidates needs exactly one staged segment; use pb2_flush");
        Segment& s = h->segs[0];
        int rc;
        if (!explicit_resident_ready(h)) {   // first call: plan building, everything in order on the handle stream
            rc = enqueue_segment(h, s, nullptr, nullptr, nullptr);
            if (rc == PB2_OK) rc = explicit_call_resident(h, s, h->stream);
            if (rc == PB2_OK) rc = amplicon_pass(h, s, h->stream);
        } else {
            rc = resident_step(h, s, true);
            if (rc != PB2_OK) return rc;
            total = (int64_t)s.h_var_count;
            if (n_records) *n_records = total;
            return PB2_OK;
        }
        if (rc == PB2_OK) rc = finish_segment(h, s);
        if (rc != PB2_OK) return rc;
        total = (int64_t)s.h_var_count;
    } else if (h->segs.size() == 1) {
        const int rc = resident_step(h, h->segs[0], false);
        if (rc != PB2_OK) return rc;
        total = (int64_t)h->segs[0].h_var_count;
    } else {
        for (auto& s : h->segs) {
            int rc = run_segment(h, s, nullptr, nullptr);
            if (rc != PB2_OK) return rc;
            total += (int64_t)s.h_var_count;
        }
    }
    if (n_records) *n_records = total;
    return PB2_OK;
}

extern "C" int pb2_resident_results(pb2_handle* h, const pb2_call_record** ref_records, const uint8_t** ref_valid, int64_t* n_loci,
                                    const pb2_call_record** variant_records, int64_t* n_variants) {
    if (!h) return PB2_ERR_ARG;
    if (h->segs.empty() || !h->segs.back().called) return fail(h, PB2_ERR_STATE, "pb2_resident_results: nothing has been called");
    const Segment& s = h->segs.back();
    if (ref_records) *ref_records = s.ref_records;
    if (ref_valid) *ref_valid = s.ref_valid;
    if (n_loci) *n_loci = s.n_loci;
    if (variant_records) *variant_records = s.var_records;
    if (n_variants) *n_variants = (int64_t)s.h_var_count;
    return PB2_OK;
}

// Flagged mismatching entries (side list of the hot kernel) against the SNV records it emitted.
//  * PB2_ENTRY_NO_CANDIDATE: the base was counted but CandidateVariantFinder never sees it ('='/'X' operations, bases past the
//    chromosome end: CandidateVariantFinder.cs:46-64,102-103) -> its support must come off the candidate.
//  * open-ended entries (only tracked when Collapse is on, RegionState.cs:114-118): VariantCollapser merges an open-ended SNV into its
//    fully anchored twin unconditionally (exact match, VariantCollapser.cs:213-215) and the smaller of two complementary open-ended
//    twins into the larger (frequency ratio >= 1 > 0.5, :218); then the record built from the total support is exact. Anything else
//    needs the explicit-candidate path.
static int reconcile_flagged_entries(pb2_handle* h, const Segment& s, const std::vector<uint32_t>& exc, std::vector<pb2_call_record>& vars) {
    struct Key { uint32_t locus; int allele; bool operator<(const Key& o) const { return locus != o.locus ? locus < o.locus : allele < o.allele; } };
    struct Acc { int nocand = 0; int open_l = 0, open_r = 0, open_lr = 0; };
    static const char base_of[4] = {'A', 'G', 'C', 'T'};
    // emitted SNV records by (position, alt) for lookup; flagged entries of alleles that were not called cannot change anything
    // (a split of a non-callable candidate is non-callable too: less support means a lower frequency and a lower q-score)
    std::unordered_map<uint64_t, const pb2_call_record*> index;   // (position, alt) is unique among SNV records
    index.reserve(vars.size() * 2);
    for (auto& v : vars)
        if (v.type == CAT_SNV) index.emplace(((uint64_t)(uint32_t)v.position << 8) | ((v.allele_bytes >> 8) & 0xff), &v);
    auto key_of = [&](uint32_t locus, int allele) {
        const int32_t pos = s.has_positions ? s.h_positions[locus] : s.first_position + (int32_t)locus;
        return ((uint64_t)(uint32_t)pos << 8) | (uint8_t)base_of[allele & 3];
    };
    // one slot per locus with a called SNV (the flagged entries are many, the called alleles few: no hashing, no sorting per entry)
    std::vector<int32_t> slot_of((size_t)s.n_loci, -1);
    int32_t n_slots = 0;
    for (auto& v : vars) {
        if (v.type != CAT_SNV) continue;
        int64_t l = -1;
        if (s.has_positions) {
            auto it = std::lower_bound(s.h_positions.begin(), s.h_positions.end(), v.position);
            if (it != s.h_positions.end() && *it == v.position) l = it - s.h_positions.begin();
        } else l = (int64_t)v.position - s.first_position;
        if (l < 0 || l >= s.n_loci) continue;
        if (slot_of[(size_t)l] < 0) slot_of[(size_t)l] = n_slots++;
    }
    std::vector<Acc> accs((size_t)n_slots * 4);
    std::vector<uint8_t> touched((size_t)n_slots * 4, 0);
    std::vector<uint32_t> slot_locus((size_t)n_slots, 0);
    for (size_t i = 0; i + 1 < exc.size(); i += 2) {
        const int allele = (int)(exc[i + 1] & 7);
        if (allele > 3 || exc[i] >= (uint32_t)s.n_loci) continue;
        const int32_t sl = slot_of[exc[i]];
        if (sl < 0) continue;
        Acc& a = accs[(size_t)sl * 4 + (size_t)allele];
        touched[(size_t)sl * 4 + (size_t)allele] = 1;
        slot_locus[(size_t)sl] = exc[i];
        const uint32_t f = exc[i + 1] & 0xffu;
        const bool l = f & PB2_ENTRY_OPEN_LEFT, r = f & PB2_ENTRY_OPEN_RIGHT;
        if (f & PB2_ENTRY_NO_CANDIDATE) a.nocand++;
        else if (l && r) a.open_lr++;
        else if (l) a.open_l++;
        else if (r) a.open_r++;
    }
    std::vector<std::pair<Key, Acc>> groups;
    for (int32_t sl = 0; sl < n_slots; sl++)
        for (int allele = 0; allele < 4; allele++)
            if (touched[(size_t)sl * 4 + (size_t)allele]) groups.push_back({Key{slot_locus[(size_t)sl], allele}, accs[(size_t)sl * 4 + (size_t)allele]});
    for (auto& g : groups) {
        auto it = index.find(key_of(g.first.locus, g.first.allele));
        if (it != index.end()) {
            const pb2_call_record& v = *it->second;
            const Acc& a = g.second;
            if (a.nocand > 0) return fail(h, PB2_ERR_UNSUPPORTED, "called SNV has support from '='/'X' operations: explicit-candidate path not built yet");
            if (!h->cfg.collapse) continue;   // open ends are not tracked without the collapser
            const int anchored = v.allele_support - (a.open_l + a.open_r + a.open_lr);
            const int kinds = (a.open_l > 0) + (a.open_r > 0) + (a.open_lr > 0);
            if (anchored > 0 || kinds <= 1) continue;                       // exact-match merge, or a single open-ended candidate on its own
            if (kinds == 2 && a.open_lr == 0) continue;                     // complementary twins: smaller merges into larger
            return fail(h, PB2_ERR_UNSUPPORTED, "open-ended SNV candidates without an anchored twin need the explicit-candidate path (not built yet)");
        }
    }
    return PB2_OK;
}

// Allele strings of a record (inline when ref_len + alt_len <= 4, else in the flush arena)
static inline void record_alleles(const pb2_call_record& r, const std::vector<uint8_t>& arena, const uint8_t*& ref, const uint8_t*& alt, uint8_t* inline_buf) {
    if (r.ref_len + r.alt_len <= 4) {
        for (int i = 0; i < 4; i++) inline_buf[i] = (uint8_t)((r.allele_bytes >> (8 * i)) & 0xff);
        ref = inline_buf;
    } else ref = arena.data() + r.allele_bytes;
    alt = ref + r.ref_len;
}
// (position, ReferenceAllele, AlternateAllele) ordinal order: SortedList by position + ComputeGenotypeAndFilterAllele's OrderBy (AlleleCaller.cs:96-140,172-176)
static bool record_less_arena(const pb2_call_record& a, const pb2_call_record& b, const std::vector<uint8_t>& arena) {
    if (a.position != b.position) return a.position < b.position;
    uint8_t ba[4], bb[4];
    const uint8_t *ra, *aa, *rb, *ab;
    record_alleles(a, arena, ra, aa, ba);
    record_alleles(b, arena, rb, ab, bb);
    auto cmp = [](const uint8_t* x, int nx, const uint8_t* y, int ny) {
        const int c = memcmp(x, y, (size_t)std::min(nx, ny));
        return c != 0 ? c : nx - ny;
    };
    const int c1 = cmp(ra, a.ref_len, rb, b.ref_len);
    if (c1 != 0) return c1 < 0;
    return cmp(aa, a.alt_len, ab, b.alt_len) < 0;
}

// Per-locus gapped-MNV reference counts of a segment (RegionState._gappedMnvReferenceCounts), or nullptr when there are none.
__global__ static void scatter_pairs_kernel(const int2* __restrict__ pairs, int32_t n, int32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicOr(out + pairs[i].x, pairs[i].y);
}
// The few loci that carry a gapped-MNV reference count or the suppress bit are scattered into a zeroed device array (no per-locus host work).
static int upload_gapped(pb2_handle* h, const Segment& s, int32_t** d_out) {
    *d_out = nullptr;
    if (h->gapped_ref.empty() && h->snv_explicit_ranges.empty()) return PB2_OK;
    std::vector<int2> pairs;
    auto locus_of = [&](int32_t pos) -> int64_t {
        if (s.has_positions) {
            auto it = std::lower_bound(s.h_positions.begin(), s.h_positions.end(), pos);
            return (it != s.h_positions.end() && *it == pos) ? (int64_t)(it - s.h_positions.begin()) : -1;
        }
        const int64_t k = (int64_t)pos - s.first_position;
        return (k >= 0 && k < s.n_loci) ? k : -1;
    };
    for (auto& kv : h->gapped_ref) {
        const int64_t l = locus_of(kv.first);
        if (l >= 0) pairs.push_back(make_int2((int)l, kv.second & (kSuppressCountSnvs - 1)));
    }
    // positions whose SNV candidates are explicit (explicit_materialize_snvs): the hot kernel must not derive SNVs from the counts there
    for (auto& r : h->snv_explicit_ranges) {
        int64_t i0, i1;   // loci with position in (r.first, r.second]
        if (s.has_positions) {
            i0 = std::upper_bound(s.h_positions.begin(), s.h_positions.end(), r.first) - s.h_positions.begin();
            i1 = std::upper_bound(s.h_positions.begin(), s.h_positions.end(), r.second) - s.h_positions.begin();
        } else {
            i0 = std::max<int64_t>(0, (int64_t)r.first + 1 - s.first_position);
            i1 = std::min<int64_t>(s.n_loci, (int64_t)r.second + 1 - s.first_position);
        }
        for (int64_t i = i0; i < i1; i++) pairs.push_back(make_int2((int)i, kSuppressCountSnvs));
    }
    if (pairs.empty()) return PB2_OK;
    int2* d_pairs = nullptr;
    CU(h, pool_alloc_t(h, d_out, (size_t)s.n_loci));
    CU(h, pool_alloc_t(h, &d_pairs, pairs.size()));
    CU(h, cudaMemsetAsync(*d_out, 0, sizeof(int32_t) * (size_t)s.n_loci, h->stream));
    CU(h, cudaMemcpyAsync(d_pairs, pairs.data(), sizeof(int2) * pairs.size(), cudaMemcpyHostToDevice, h->stream));
    scatter_pairs_kernel<<<(unsigned)((pairs.size() + 255) / 256), 256, 0, h->stream>>>(d_pairs, (int32_t)pairs.size(), *d_out);
    CU(h, cudaStreamSynchronize(h->stream));   // pairs is a local
    pool_free(h, d_pairs);
    return PB2_OK;
}

// The explicit-candidate batches of this flush, replaying the batches SmallVariantCaller.Execute would have formed (SmallVariantCaller.cs:88-104,
// RegionStateManager.GetCandidatesToProcess :283-334): a batch happens whenever upTo = (read position - 1) enters a new 1000-bp block key; it takes
// the existing blocks that end at or before upTo, in order, up to the first block holding an allele that extends past upTo, plus — when an included
// allele reaches past the last included block — the finished open-left SNV/MNV candidates of later blocks (AddCollapsableFromOtherBlocks :441-457).
// Returns the last position cleared (0 = none; INT32_MAX = everything).
static int run_explicit_batches(pb2_handle* h, int32_t up_to, bool reads_path, std::vector<pb2_call_record>& called, std::vector<pb2_call_record_ext>& called_ext,
                                int32_t* cleared_out) {
    *cleared_out = h->cleared_through;
    auto all_alive_sorted = [&]() {
        std::vector<size_t> idx;
        for (size_t i = 0; i < h->cands.size(); i++) if (h->cands[i].alive) idx.push_back(i);
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return h->cands[a].position < h->cands[b].position; });
        return idx;
    };
    if (!reads_path) {   // locus-major pushes are complete by construction: one batch, nothing left uncleared
        explicit_add_forced_candidates(h, -1);
        const int rc = explicit_call_batch(h, all_alive_sorted(), -1, 0, INT32_MAX, called, called_ext);
        *cleared_out = INT32_MAX;
        return rc;
    }
    // blocks that exist: touched by a kept read, or holding a candidate
    std::vector<int32_t> keys;
    {
        const DeviceReads& R = h->reads;
        if (R.n > 0) {
            const int32_t a = std::max(R.min_start, h->cleared_through + 1), e = R.max_end;
            if (a <= e) {
                const int32_t k0 = (a + 999) / 1000, nk = (e + 999) / 1000 - k0 + 1;
                std::vector<uint32_t> bits((size_t)(nk + 31) / 32, 0);
                uint32_t* d_bits = nullptr;
                CU(h, pool_alloc_t(h, &d_bits, bits.size()));
                CU(h, cudaMemsetAsync(d_bits, 0, sizeof(uint32_t) * bits.size(), h->stream));
                CU(h, launch_reads_block_bitmap(R.pos0.p, R.end_pos.p, R.n, h->cleared_through, k0, nk, d_bits, h->stream));
                CU(h, cudaMemcpyAsync(bits.data(), d_bits, sizeof(uint32_t) * bits.size(), cudaMemcpyDeviceToHost, h->stream));
                CU(h, cudaStreamSynchronize(h->stream));
                pool_free(h, d_bits);
                for (int32_t k = 0; k < nk; k++) if ((bits[(size_t)k >> 5] >> (k & 31)) & 1u) keys.push_back(k0 + k);
            }
        }
        for (auto& c : h->cands) if (c.alive) keys.push_back((c.position + 999) / 1000);
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    }
    // the alive candidates in position order, computed once and again only when a batch returned candidates to the state
    std::vector<size_t> order = all_alive_sorted();
    size_t order_cands = h->cands.size(), cursor = 0;
    auto refresh_order = [&]() {
        if (h->cands.size() != order_cands) { order = all_alive_sorted(); order_cands = h->cands.size(); cursor = 0; }
    };
    const bool can_defer = h->forced.empty() && h->forced_positions.empty() && h->forced_pending.empty();
    std::vector<size_t> deferred;
    auto flush_deferred = [&]() -> int {
        if (deferred.empty()) return PB2_OK;
        std::vector<size_t> none;
        const int rc = explicit_call_batch(h, deferred, -1, 0, INT32_MAX, called, called_ext, &none);
        deferred.clear();
        return rc;
    };
    Trace tr("explicit_batches");
    tr.mark("keys");
    extern double g_batch_phase_ms[8];
    for (double& v : g_batch_phase_ms) v = 0;
    struct PhaseReport {
        ~PhaseReport() {
            if (!trace_on()) return;
            fprintf(stderr, "[pb2] batch phases: copy=%.1f collapse_score=%.1f collapse=%.1f mnv_score=%.1f refs=%.1f realloc=%.1f final_score=%.1f emit=%.1f ms\n", g_batch_phase_ms[0],
                    g_batch_phase_ms[1], g_batch_phase_ms[2], g_batch_phase_ms[3], g_batch_phase_ms[4], g_batch_phase_ms[5], g_batch_phase_ms[6], g_batch_phase_ms[7]);
        }
    } phase_report;
    std::vector<int32_t> fire;   // upTo values in call order; -1 = the final Call(null)
    {
        size_t used = 0;
        for (int32_t t : h->triggers) { if (up_to >= 0 && t > up_to) break; fire.push_back(t); used++; }
        h->triggers.erase(h->triggers.begin(), h->triggers.begin() + (long)used);
        fire.push_back(up_to >= 0 ? up_to : -1);
    }
    for (int32_t t : fire) {
        if (!h->forced_pending.empty()) {   // AddForcedAlleleAsCandidate(upTo) runs before Call(upTo) (SmallVariantCaller.cs:99-104); AddCandidates creates the block
            for (auto& kv : h->forced_pending) { if (t >= 0 && kv.first > t) break; keys.push_back((kv.first + 999) / 1000); }
            std::sort(keys.begin(), keys.end());
            keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
            explicit_add_forced_candidates(h, t);
        }
        const int key = t < 0 ? -1 : (t <= 0 ? 0 : (t + 999) / 1000);
        if (t >= 0 && key == h->last_trigger_key) continue;     // GetCandidatesToProcess returns null (:288-291)
        h->last_trigger_key = key;
        int32_t max_end = 0, max_endpoint = 0;
        for (int32_t k : keys) {
            if ((int64_t)k * 1000 <= *cleared_out) continue;                    // already processed and recycled
            if (t >= 0 && !((int64_t)k * 1000 <= t)) continue;
            auto it = h->block_max_endpoint.find(k);
            const int32_t mep = it == h->block_max_endpoint.end() ? 0 : it->second;
            if (t >= 0 && mep > t) break;                                        // :305-306
            max_end = k * 1000;
            max_endpoint = std::max(max_endpoint, mep);
        }
        if (max_end == 0) continue;
        const bool pull_collapsable = t >= 0 && max_endpoint > max_end && h->cfg.collapse;   // AddCollapsableFromOtherBlocks (:441-457)
        if (pull_collapsable && !h->cfg.call_mnvs) {
            // the count-based SNVs of the later blocks' positions <= upTo become explicit candidates first (see explicit_materialize_snvs)
            int32_t lo = max_end;
            for (auto& r : h->snv_explicit_ranges) lo = std::max(lo, r.second);
            const int rc = explicit_materialize_snvs(h, lo, t);
            if (rc != PB2_OK) return rc;
        }
        std::vector<size_t> batch, kill;
        refresh_order();
        while (cursor < order.size() && h->cands[order[cursor]].position <= *cleared_out) cursor++;
        for (size_t k = cursor; k < order.size() && h->cands[order[k]].position <= max_end; k++)
            if (h->cands[order[k]].alive) { batch.push_back(order[k]); kill.push_back(order[k]); }
        if (!pull_collapsable && can_defer) {
            // A batch of fully anchored insertions / deletions only: every allele is scored on its own (no collapsing, no MNV reallocation, no
            // gapped-MNV reference counts, nothing returns to the state), so consecutive batches of this kind are scored in one device pass
            bool boring = true;
            for (size_t i : batch) {
                const HostCand& c = h->cands[i];
                if ((c.type != CAT_INS && c.type != CAT_DEL) || (h->cfg.collapse && (c.open_left || c.open_right))) { boring = false; break; }
            }
            if (boring) {
                for (size_t i : batch) { h->cands[i].alive = false; deferred.push_back(i); }
                *cleared_out = t >= 0 ? max_end : INT32_MAX;
                continue;
            }
        }
        { const int rcd = flush_deferred(); if (rcd != PB2_OK) return rcd; }
        if (pull_collapsable) {
            // ExtractCollapsable(upTo) of the blocks that start after max_end and at or before upTo (RegionState.cs:470-490), position by position; each
            // collapsable goes to the batch, and List.Remove takes the FIRST candidate of the position's list that Equals it out of the state
            const std::vector<size_t> order = all_alive_sorted();
            for (size_t a = 0; a < order.size();) {
                size_t b = a;
                while (b < order.size() && h->cands[order[b]].position == h->cands[order[a]].position) b++;
                const HostCand& first = h->cands[order[a]];
                const int32_t block_start = ((first.position + 999) / 1000 - 1) * 1000 + 1;
                if (first.position > max_end && block_start <= t) {
                    std::vector<size_t> list(order.begin() + (long)a, order.begin() + (long)b);
                    std::vector<size_t> collapsables;
                    for (size_t i : list) {
                        const HostCand& c = h->cands[i];
                        if (c.position + (int)c.alt.size() - 1 <= t && !c.open_right && (c.type == CAT_MNV || c.type == CAT_SNV)) collapsables.push_back(i);
                    }
                    for (size_t c : collapsables) {
                        batch.push_back(c);
                        for (size_t k = 0; k < list.size(); k++)
                            if (h->cands[list[k]].Equals(h->cands[c])) { kill.push_back(list[k]); list.erase(list.begin() + (long)k); break; }
                    }
                }
                a = b;
            }
        }
        const int rc = explicit_call_batch(h, batch, t >= 0 ? max_end : -1, *cleared_out == INT32_MAX ? INT32_MAX : *cleared_out, t >= 0 ? max_end : INT32_MAX, called,
                                           called_ext, &kill);
        if (rc != PB2_OK) return rc;
        *cleared_out = t >= 0 ? max_end : INT32_MAX;
    }
    tr.mark("loop");
    const int rcf = flush_deferred();
    tr.mark("deferred");
    return rcf;
}

// The per-locus part of AlleleCaller.ComputeGenotypeAndFilterAllele (:143-177) for the germline genotypers, over the records of one flush (already
// grouped by position, reference alleles pruned where a variant was called): DiploidThresholdingGenotyper.SetGenotypes / HaploidGenotyper.SetGenotypes
// with GenotypeCalculatorUtilities (the alleles of a locus ordered by frequency, 0.20 / 0.70 / 0.80 thresholds, tri-allelic check, pruning) and then
// DiploidLocusProcessor.Process (DiploidLocusProcessor.cs:13-51). A locus that only holds its reference allele was genotyped on the device.
static void germline_locus_pass(pb2_handle* h) {
    const DeviceConfig& d = h->dcfg;
    const bool germline = d.ploidy != PLOIDY_SOMATIC;
    const bool locus_processor = h->cfg.ploidy == PLOIDY_DIPLOID;   // follows the SAMPLE ploidy (Factory.cs:145-147)
    if (!germline && !locus_processor) return;
    auto& R = h->h_out;
    auto& E = h->h_out_ext;
    auto freq_of = [](int support, int total) { return total == 0 ? 0.0f : std::min((float)support / (float)total, 1.0f); };
    std::vector<uint8_t> drop(R.size(), 0);
    bool any_drop = false;
    for (size_t b = 0; b < R.size();) {
        size_t e = b + 1;
        while (e < R.size() && R[e].position == R[b].position) e++;
        std::vector<size_t> ng;   // allelesAtPosition.Where(x => !x.IsForcedToReport)
        bool any_variant = false;
        for (size_t i = b; i < e; i++) if (!(R[i].sb_flags & 8)) { ng.push_back(i); any_variant |= R[i].type != CAT_REF; }
        if (germline && any_variant) {
            const bool hap = d.ploidy == PLOIDY_HAPLOID;
            const float minor = d.diploid_minor_vf, major = d.diploid_major_vf, sum_vf = d.diploid_sum_vf;
            // FilterAndOrderAllelesByFrequency (GenotypeCalculatorUtilities.cs:60-82): ng is in (ref, alt) order already, the sort is stable
            std::vector<size_t> ordered, prune;
            for (size_t i : ng) {
                if (R[i].type == CAT_REF) continue;
                if ((double)freq_of(R[i].allele_support, R[i].total_coverage) >= (double)minor) ordered.push_back(i); else prune.push_back(i);
            }
            std::stable_sort(ordered.begin(), ordered.end(), [&](size_t x, size_t y) {
                return freq_of(R[x].allele_support, R[x].total_coverage) > freq_of(R[y].allele_support, R[y].total_coverage);
            });
            // GetReferenceFrequency (:85-130)
            double reference_frequency = 0;
            if (ng.size() == 1) reference_frequency = freq_of(R[ng[0]].reference_support, R[ng[0]].total_coverage);
            else {
                double by_snp = 0, indel = 0;
                bool returned = false;
                for (size_t i : ng) {
                    if (R[i].type == CAT_REF) { reference_frequency = freq_of(R[i].allele_support, R[i].total_coverage); returned = true; break; }
                    if (R[i].type == CAT_SNV) by_snp = freq_of(R[i].reference_support, R[i].total_coverage);
                    else indel += freq_of(R[i].allele_support, R[i].total_coverage);
                }
                if (!returned) reference_frequency = std::max(by_snp - indel, 0.0);
            }
            const bool ref_exists = reference_frequency >= (double)minor;
            bool depth_issue = false;
            for (size_t i : ng) depth_issue |= R[i].total_coverage < d.min_coverage;
            const float top = ordered.empty() ? 0.0f : freq_of(R[ordered[0]].allele_support, R[ordered[0]].total_coverage);
            const bool ref_call = ordered.empty() || top < minor;
            int gt;
            bool multi_allelic = false;
            if (hap) {   // HaploidGenotyper.CalculateHaploidGenotype (HaploidGenotyper.cs:54-82)
                gt = GT_HEMI_NOCALL;
                if (!depth_issue && ref_call && ref_exists && reference_frequency > (double)major) gt = GT_HEMI_REF;
                if (!depth_issue && !ref_call && !ref_exists && top > major) gt = GT_HEMI_ALT;
            } else {     // CalculateDiploidGenotype (DiploidThresholdingGenotyper.cs:77-125) + ConvertSimpleGenotypeToComplexGenotype (:160-233)
                enum { P_HOM_REF, P_HET, P_HOM_ALT } prelim;
                if (ref_call) prelim = P_HOM_REF;
                else if (top >= minor && top <= major) prelim = P_HET;
                else if (top > major) prelim = P_HOM_ALT;
                else prelim = P_HOM_REF;
                if (depth_issue) gt = ref_call ? GT_REF_NOCALL : GT_ALT_NOCALL;
                else if (prelim == P_HOM_REF) {
                    const pb2_call_record& first = R[ng[0]];
                    if (!ref_exists) gt = GT_REF_NOCALL;
                    else if (first.type == CAT_REF && (1 - freq_of(first.allele_support, first.total_coverage)) > minor) gt = GT_REF_AND_NOCALL;
                    else gt = GT_HOM_REF;
                } else if (prelim == P_HET) {
                    if (ordered.size() == 1) gt = ref_exists ? GT_HET_ALT_REF : GT_ALT_AND_NOCALL;
                    else {
                        // CheckForTriAllelicIssue (:133-150)
                        bool fail = false;
                        if (R[ordered.back()].type == CAT_SNV) {
                            const float f0 = top, f1 = freq_of(R[ordered[1]].allele_support, R[ordered[1]].total_coverage);
                            if (ref_exists && (((double)f0 + reference_frequency) < (double)sum_vf)) fail = true;
                            else fail = (f0 + f1) < sum_vf;
                        }
                        if (fail) { multi_allelic = true; gt = ref_exists ? GT_ALT_NOCALL : GT_ALT12_NOCALL; }
                        else gt = ref_exists ? GT_HET_ALT_REF : GT_HET_ALT12;
                    }
                } else gt = GT_HOM_ALT;
            }
            // GetAllelesToPruneBasedOnGTCall (:11-46)
            int allowed = 0;
            if (gt == GT_ALT_AND_NOCALL || gt == GT_ALT_NOCALL || gt == GT_HOM_ALT || gt == GT_HET_ALT_REF || gt == GT_HEMI_ALT) allowed = 1;
            else if (gt == GT_ALT12_NOCALL || gt == GT_HET_ALT12) allowed = 2;
            for (size_t k = 0; k < ordered.size(); k++) if ((int)k >= allowed) prune.push_back(ordered[k]);
            int phase_set_index = 1;   // DiploidThresholdingGenotyper.SetGenotypes (:57-72); the haploid genotyper leaves it unset
            for (size_t i : ng) {
                if (!hap) E[i].phase_set_index = R[i].type == CAT_REF ? 0 : phase_set_index++;
                R[i].genotype = (uint8_t)gt;
                R[i].genotype_qscore = germline_gq(hap, gt, R[i].total_coverage, R[i].allele_support, d.min_gq, d.max_gq);
                if (multi_allelic) R[i].filters |= (uint16_t)(1u << FLT_MULTI_ALLELIC);
                R[i].filters &= (uint16_t)~(1u << FLT_LOW_GQ);
                if (d.low_gq_filter >= 0 && (float)R[i].genotype_qscore < (float)d.low_gq_filter) R[i].filters |= (uint16_t)(1u << FLT_LOW_GQ);
            }
            for (size_t i : prune) {   // pruned unless it is one of the forced alleles (AlleleCaller.cs:153-162)
                uint8_t buf[4];
                const uint8_t *ra, *aa;
                record_alleles(R[i], h->arena, ra, aa, buf);
                if (!h->forced.empty() && h->forced.count(std::make_tuple(R[i].position, std::string((const char*)ra, R[i].ref_len), std::string((const char*)aa, R[i].alt_len))))
                    continue;
                drop[i] = 1;
                any_drop = true;
            }
        }
        if (locus_processor) {   // DiploidLocusProcessor.Process
            std::vector<size_t> forced, non_forced;
            for (size_t i = b; i < e; i++) if (!drop[i]) ((R[i].filters >> FLT_FORCED_REPORT) & 1 ? forced : non_forced).push_back(i);
            if (!forced.empty()) {
                bool is_ref = false, any_nocall = false;
                int min_gq = 0;
                for (size_t k = 0; k < non_forced.size(); k++) {
                    const pb2_call_record& r = R[non_forced[k]];
                    is_ref |= r.type == CAT_REF;
                    any_nocall |= r.genotype == GT_ALT12_NOCALL || r.genotype == GT_ALT_NOCALL || r.genotype == GT_HEMI_NOCALL || r.genotype == GT_REF_NOCALL;
                    min_gq = k == 0 ? r.genotype_qscore : std::min(min_gq, r.genotype_qscore);
                }
                const int gt = (non_forced.empty() || any_nocall) ? GT_ALT_NOCALL : (is_ref ? GT_HOM_REF : GT_OTHERS);
                for (size_t i : forced) R[i].genotype = (uint8_t)gt;
                for (size_t i = b; i < e; i++) if (!drop[i]) R[i].genotype_qscore = min_gq;
            }
        }
        b = e;
    }
    if (any_drop) {
        size_t w = 0;
        for (size_t i = 0; i < R.size(); i++) if (!drop[i]) { R[w] = R[i]; E[w] = E[i]; w++; }
        R.resize(w); E.resize(w);
    }
}

// After a flush that cleared positions <= cleared_to: the reads that end inside the cleared positions leave the store (device compaction).
static int compact_reads(pb2_handle* h, int32_t cleared_to) {
    DeviceReads& R = h->reads;
    if (R.n == 0) return PB2_OK;
    cudaStream_t st = h->stream;
    const size_t n1 = (size_t)R.n + 1;
    int64_t *flags = nullptr, *lc = nullptr, *ls = nullptr, *ni = nullptr, *nc = nullptr, *ns = nullptr;
    void* temp = nullptr;
    CU(h, pool_alloc_t(h, &flags, n1)); CU(h, pool_alloc_t(h, &lc, n1)); CU(h, pool_alloc_t(h, &ls, n1));
    CU(h, pool_alloc_t(h, &ni, n1)); CU(h, pool_alloc_t(h, &nc, n1)); CU(h, pool_alloc_t(h, &ns, n1));
    CU(h, launch_reads_keep_flags(R.end_pos.p, R.n, cleared_to, R.cigar_off.p, R.seq_off.p, flags, lc, ls, st));
    size_t tb = 0;
    CU(h, exclusive_scan_i64(flags, ni, (int64_t)n1, nullptr, 0, &tb, st));
    CU(h, pool_alloc(h, &temp, std::max<size_t>(tb, 16)));
    CU(h, exclusive_scan_i64(flags, ni, (int64_t)n1, temp, tb, nullptr, st));
    CU(h, exclusive_scan_i64(lc, nc, (int64_t)n1, temp, tb, nullptr, st));
    CU(h, exclusive_scan_i64(ls, ns, (int64_t)n1, temp, tb, nullptr, st));
    int64_t totals[3] = {0, 0, 0};
    CU(h, cudaMemcpyAsync(&totals[0], ni + R.n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(&totals[1], nc + R.n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(&totals[2], ns + R.n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    if (totals[0] == 0) {
        const int32_t last = R.last_pos0;
        free_reads(h);
        h->reads.last_pos0 = last;
    } else if (totals[0] < R.n) {
        DeviceReads K;
        K.has_dirs = R.has_dirs; K.has_collapsed = R.has_collapsed; K.has_amplicon = R.has_amplicon;
        const size_t kn = (size_t)totals[0], kc = (size_t)totals[1], ks = (size_t)totals[2];
        CU(h, grow(h, K.pos0, kn, 0)); CU(h, grow(h, K.end_pos, kn, 0)); CU(h, grow(h, K.flag, kn, 0));
        CU(h, grow(h, K.cigar_off, kn + 1, 0)); CU(h, grow(h, K.seq_off, kn + 1, 0));
        CU(h, grow(h, K.cigar, std::max<size_t>(kc, 1), 0)); CU(h, grow(h, K.bases, std::max<size_t>(ks, 1), 0)); CU(h, grow(h, K.quals, std::max<size_t>(ks, 1), 0));
        if (R.has_dirs) CU(h, grow(h, K.base_dirs, std::max<size_t>(ks, 1), 0));
        if (R.has_collapsed) CU(h, grow(h, K.collapsed, kn, 0));
        if (R.has_amplicon) CU(h, grow(h, K.amplicon, kn, 0));
        CU(h, grow(h, K.slots, ks + 32, 0));
        ReadsCompactArgs a;
        a.n = R.n; a.new_index = ni; a.new_cigar = nc; a.new_seq = ns;
        a.pos0 = R.pos0.p; a.end_pos = R.end_pos.p; a.flag = R.flag.p; a.cigar_off = R.cigar_off.p; a.cigar = R.cigar.p; a.seq_off = R.seq_off.p;
        a.bases = R.bases.p; a.quals = R.quals.p; a.base_dirs = R.has_dirs ? R.base_dirs.p : nullptr; a.collapsed = R.has_collapsed ? R.collapsed.p : nullptr;
        a.o_pos0 = K.pos0.p; a.o_end_pos = K.end_pos.p; a.o_flag = K.flag.p; a.o_cigar_off = K.cigar_off.p; a.o_cigar = K.cigar.p; a.o_seq_off = K.seq_off.p;
        a.o_bases = K.bases.p; a.o_quals = K.quals.p; a.o_base_dirs = K.base_dirs.p; a.o_collapsed = K.collapsed.p;
        a.slots = R.slots.p + 16; a.o_slots = K.slots.p + 16;
        a.amplicon = R.has_amplicon ? R.amplicon.p : nullptr; a.o_amplicon = K.amplicon.p;
        CU(h, launch_reads_compact(a, st));
        K.n = totals[0]; K.n_cigar = totals[1]; K.n_seq = totals[2];
        K.min_start = R.min_start; K.max_end = R.max_end; K.last_pos0 = R.last_pos0;
        CU(h, cudaStreamSynchronize(st));
        free_reads(h);
        h->reads = K;
    }
    void* ptrs[] = {flags, lc, ls, ni, nc, ns, temp};
    for (void* p : ptrs) pool_free(h, p);
    return PB2_OK;
}

template <class T>
struct HostSpan {
    T* p = nullptr;
    size_t n = 0;
    bool empty() const { return n == 0; }
    T& operator[](size_t i) const { return p[i]; }
};
static int flush_impl(pb2_handle* h, int32_t up_to_position, const pb2_call_record** out, int64_t* n, bool keep_reads);
extern "C" int pb2_flush(pb2_handle* h, int32_t up_to_position, const pb2_call_record** out, int64_t* n) { return flush_impl(h, up_to_position, out, n, false); }
// IAlleleCaller.Call for everything staged through pb2_push_reads, the reads staying on the device: the candidates are found again in the stored reads
// (CandidateVariantFinder.FindCandidates), the batches of SmallVariantCaller.Execute replayed again, and nothing is consumed - the whole job from
// device-resident reads, repeatable (bench.py, and hosts that re-call a chromosome with other options' worth of candidates).
extern "C" int pb2_flush_resident(pb2_handle* h, const pb2_call_record** out, int64_t* n) {
    if (!h || !out || !n) return fail(h, PB2_ERR_ARG, "pb2_flush_resident: null argument");
    if (h->cleared_through != 0) return fail(h, PB2_ERR_STATE, "pb2_flush_resident: positions were already cleared by a partial pb2_flush");
    CU(h, cudaSetDevice(h->device));
    const std::vector<int32_t> triggers = h->triggers;
    const int32_t push_last_key = h->push_last_key;
    h->cands.clear(); h->cand_by_pos.clear(); h->block_max_endpoint.clear(); h->gapped_ref.clear(); h->snv_explicit_ranges.clear();
    h->last_trigger_key = 0;
    rearm_forced(h);
    int rc = explicit_find_candidates(h, 0, nullptr);
    if (rc == PB2_OK) rc = flush_impl(h, -1, out, n, true);
    h->triggers = triggers;
    h->push_last_key = push_last_key;
    return rc;
}
static int flush_impl(pb2_handle* h, int32_t up_to_position, const pb2_call_record** out, int64_t* n, bool keep_reads) {
    if (!h || !out || !n) return fail(h, PB2_ERR_ARG, "pb2_flush: null argument");
    CU(h, cudaSetDevice(h->device));
    nvtx_range nv("pb2_flush");
    Trace tr("flush");
    h->h_out.clear();
    h->arena.clear();
    const bool reads_path = h->reads.size() != 0 || !h->triggers.empty();
    for (size_t i = 0; i < h->segs.size();) {   // a resident staging of the reads (pb2_stage_reads) is superseded: the flush stages what it needs itself
        if (h->segs[i].from_reads) { free_segment(h, h->segs[i]); h->segs.erase(h->segs.begin() + (long)i); explicit_release_resident(h); release_resident_graph(h); } else i++;
    }
    if (h->resident_explicit) for (auto& s : h->segs) s.called = false;   // a pb2_call_resident pass appended explicit alleles to the variant stream: redo
    // reads staged through pb2_push_reads: positions in complete 1000-bp blocks <= up_to_position are callable (RegionStateManager.cs:283-314);
    // locus-major pushes are complete by construction
    const int32_t cleared_end = up_to_position < 0 ? INT32_MAX : (up_to_position / 1000) * 1000;
    {
        // counts are final for every position <= up_to (reads arrive in position order): stage them all, so that alleles reaching past the last
        // complete block find their end-point coverage; only positions inside cleared blocks are emitted below
        const int rc = stage_reads_segment(h, up_to_position < 0 ? INT32_MAX : up_to_position, h->cleared_through + 1);
        if (rc != PB2_OK) return rc;
    }
    tr.mark("stage_reads");
    // explicit candidates first: their gapped-MNV reference counts feed the point alleles of the hot kernel
    std::vector<pb2_call_record> explicit_called;
    std::vector<pb2_call_record_ext> explicit_ext;
    int32_t cleared_to = cleared_end;
    {
        const int rc = run_explicit_batches(h, up_to_position, reads_path, explicit_called, explicit_ext, &cleared_to);
        if (rc != PB2_OK) return rc;
        if (!reads_path) cleared_to = INT32_MAX;
    }
    tr.mark("explicit");
    struct OutRec { pb2_call_record r; pb2_call_record_ext e; };
    pb2_call_record_ext zero_ext;
    memset(&zero_ext, 0, sizeof(zero_ext));
    h->h_out_ext.clear();
    auto emit = [&](const OutRec& o) {
        if (h->dcfg.own_hi > 0 && (o.r.position < h->dcfg.own_lo || o.r.position > h->dcfg.own_hi)) return;   // an interval shard emits the positions it owns
        h->h_out.push_back(o.r); h->h_out_ext.push_back(o.e);
    };
    std::vector<uint8_t> explicit_used(explicit_called.size(), 0);
    const bool want_collapsed = h->cfg.expect_collapsed != 0;
    for (auto& s : h->segs) {
        int32_t* d_collapsed = nullptr;   // ReadCollapsedCountTotal source: CollapsedRegionState._collapsedCount per locus
        std::vector<int32_t> collapsed;
        if (want_collapsed) {
            CU(h, pool_alloc_t(h, &d_collapsed, (size_t)s.n_loci * kNumCollapsed));
            s.called = false;
        }
        if (!s.called) {
            int32_t* d_gapped = nullptr;
            int rc = upload_gapped(h, s, &d_gapped);
            if (rc == PB2_OK) rc = run_segment(h, s, nullptr, d_collapsed, d_gapped);
            pool_free(h, d_gapped);
            if (rc != PB2_OK) return rc;
        }
        std::vector<pb2_call_record> hot_vars((size_t)s.h_var_count);
        std::vector<uint32_t> exc((size_t)s.h_exc_count * 2);
        if (!exc.empty()) CU(h, cudaMemcpyAsync(exc.data(), s.exc_entries, sizeof(uint32_t) * exc.size(), cudaMemcpyDeviceToHost, h->stream));
        if (!hot_vars.empty()) CU(h, cudaMemcpyAsync(hot_vars.data(), s.var_records, sizeof(pb2_call_record) * hot_vars.size(), cudaMemcpyDeviceToHost, h->stream));
        // the dense reference stream lands in pinned memory the handle keeps (96 MB per million loci: a pageable vector made this copy the flush's
        // largest item, at a tenth of the link's rate and with a page fault per 4 KB)
        HostSpan<pb2_call_record> refs;
        HostSpan<uint8_t> valid;
        if (s.ref_records) {
            auto pinned = [&](void*& p, size_t& have, size_t need) -> cudaError_t {
                if (need <= have) return cudaSuccess;
                if (p) cudaFreeHost(p);
                p = nullptr; have = 0;
                const cudaError_t e = cudaHostAlloc(&p, need + need / 4, cudaHostAllocDefault);
                if (e == cudaSuccess) have = need + need / 4;
                return e;
            };
            CU(h, pinned(h->pin_refs, h->pin_refs_bytes, sizeof(pb2_call_record) * (size_t)s.n_loci));
            CU(h, pinned(h->pin_valid, h->pin_valid_bytes, (size_t)s.n_loci));
            refs = HostSpan<pb2_call_record>{static_cast<pb2_call_record*>(h->pin_refs), (size_t)s.n_loci};
            valid = HostSpan<uint8_t>{static_cast<uint8_t*>(h->pin_valid), (size_t)s.n_loci};
            CU(h, cudaMemcpyAsync(refs.p, s.ref_records, sizeof(pb2_call_record) * refs.n, cudaMemcpyDeviceToHost, h->stream));
            CU(h, cudaMemcpyAsync(valid.p, s.ref_valid, valid.n, cudaMemcpyDeviceToHost, h->stream));
        }
        if (d_collapsed) {
            collapsed.resize((size_t)s.n_loci * kNumCollapsed);
            CU(h, cudaMemcpyAsync(collapsed.data(), d_collapsed, sizeof(int32_t) * collapsed.size(), cudaMemcpyDeviceToHost, h->stream));
        }
        CU(h, cudaStreamSynchronize(h->stream));
        pool_free(h, d_collapsed);
        tr.mark("hot+d2h");
        if (!exc.empty() && !h->cfg.call_mnvs) { const int rc = reconcile_flagged_entries(h, s, exc, hot_vars); if (rc != PB2_OK) return rc; }
        tr.mark("reconcile");
        auto locus_of = [&](int32_t pos) -> int64_t {
            if (s.has_positions) {
                auto it = std::lower_bound(s.h_positions.begin(), s.h_positions.end(), pos);
                return (it != s.h_positions.end() && *it == pos) ? (int64_t)(it - s.h_positions.begin()) : -1;
            }
            const int64_t k = (int64_t)pos - s.first_position;
            return (k >= 0 && k < s.n_loci) ? k : -1;
        };
        auto with_totals = [&](OutRec o) {   // CollapsedCoverageCalculator (:18-37): the counts at the allele's position / spanning start point
            const int times = o.e.collapsed_total[0] == kCollapsedTotalTwice ? 2 : 1;   // MNV candidates: see explicit_call_batch
            if (times == 2) o.e.collapsed_total[0] = 0;
            if (!collapsed.empty()) {
                const int64_t l = locus_of(o.r.type == CAT_DEL ? o.r.position + 1 : o.r.position);
                if (l >= 0) for (int t = 0; t < kNumCollapsed; t++) o.e.collapsed_total[t] = times * collapsed[(size_t)l * kNumCollapsed + t];
            }
            return o;
        };
        std::vector<OutRec> vars;
        vars.reserve(hot_vars.size());
        for (auto& r : hot_vars) vars.push_back(with_totals(OutRec{r, zero_ext}));
        // the explicit alleles called inside this segment's positions join its variant stream
        const int32_t seg_lo = s.has_positions ? (s.h_positions.empty() ? 1 : s.h_positions.front()) : s.first_position;
        const int32_t seg_hi = s.has_positions ? (s.h_positions.empty() ? 0 : s.h_positions.back()) : (int32_t)(s.first_position + s.n_loci - 1);
        std::multimap<int32_t, OutRec> ref_override;   // reference alleles that gained support from a reallocated MNV (MnvReallocator.cs:255-265)
        for (size_t k = 0; k < explicit_called.size(); k++)
            if (!explicit_used[k] && explicit_called[k].position >= seg_lo && explicit_called[k].position <= seg_hi) {
                const OutRec o = with_totals(OutRec{explicit_called[k], explicit_ext[k]});
                explicit_used[k] = 1;
                // CallMNVs off: a forced SNV that is callable on its own merits is already in the hot kernel's variant stream
                if (!h->cfg.call_mnvs && o.r.type == CAT_SNV && !(o.r.sb_flags & 8)) {
                    bool explicit_here = false;   // ... unless the SNV candidates of this position are explicit (the hot kernel derives none there)
                    for (auto& r : h->snv_explicit_ranges) explicit_here |= o.r.position > r.first && o.r.position <= r.second;
                    if (!explicit_here) continue;
                }
                if (o.r.type == CAT_REF) ref_override.insert({o.r.position, o});
                else vars.push_back(o);
            }
        if (!h->forced.empty()) {
            // AlleleCaller.ComputeGenotypeAndFilterAllele (:143-150): an allele that is only reported because it is forced (IsForcedToReport) does not prune
            // the reference allele of its position; the reference allele then takes its place among them in (ref, alt) order
            std::set<int32_t> real_pos, forced_pos;
            for (auto& v : vars) ((v.r.sb_flags & 8) ? forced_pos : real_pos).insert(v.r.position);
            for (int32_t pos : forced_pos) {
                const int64_t l = locus_of(pos);
                if (!real_pos.count(pos)) {
                    auto ov = ref_override.equal_range(pos);
                    if (ov.first != ov.second) for (auto it = ov.first; it != ov.second; ++it) vars.push_back(it->second);
                    else if (!refs.empty() && l >= 0 && valid[(size_t)l]) vars.push_back(with_totals(OutRec{refs[(size_t)l], zero_ext}));
                }
                ref_override.erase(pos);
                if (!valid.empty() && l >= 0) valid[(size_t)l] = 0;
            }
            if (!h->cfg.output_gvcf) {   // reference candidates of forced positions (RegionState.GetAllCandidates :393-449) where no forced allele was left to report
                for (auto& kv : ref_override) if (!real_pos.count(kv.first)) vars.push_back(kv.second);
                ref_override.clear();
            }
        }
        {   // (position, ref, alt) order: sort 8-byte keys by position, then settle the few positions that hold several alleles
            std::vector<uint64_t> keys(vars.size());
            for (size_t k = 0; k < vars.size(); k++) keys[k] = ((uint64_t)(uint32_t)vars[k].r.position << 32) | (uint32_t)k;
            std::sort(keys.begin(), keys.end());
            for (size_t a = 0; a < keys.size();) {
                size_t b = a + 1;
                while (b < keys.size() && (keys[b] >> 32) == (keys[a] >> 32)) b++;
                if (b - a > 1)
                    std::stable_sort(keys.begin() + (long)a, keys.begin() + (long)b, [&](uint64_t x, uint64_t y) {
                        return record_less_arena(vars[(uint32_t)x].r, vars[(uint32_t)y].r, h->arena);
                    });
                a = b;
            }
            std::vector<OutRec> sorted;
            sorted.reserve(vars.size());
            for (uint64_t k : keys) sorted.push_back(vars[(uint32_t)k]);
            vars.swap(sorted);
        }
        // merge the dense reference stream (already in position order) with the sorted variant stream; a reference allele is pruned wherever a
        // variant was called (AlleleCaller.cs:146-147)
        size_t vi = 0;
        if (refs.empty() && ref_override.empty()) {   // no reference stream: the sorted variant stream is the output
            while (vi < vars.size() && vars[vi].r.position <= cleared_to) emit(vars[vi++]);
            continue;
        }
        // Two passes: the walk decides what goes where (a 4-byte plan entry per output record: reference locus i, or variant k), the fill copies the
        // 96 + 80 bytes per record with a few host threads - a gVCF chromosome is a million reference records per million loci, and one thread
        // writes them at ~6 GB/s.
        {
            const bool owned = h->dcfg.own_hi > 0;
            auto mine = [&](int32_t pos) { return !owned || (pos >= h->dcfg.own_lo && pos <= h->dcfg.own_hi); };
            constexpr uint32_t kVar = 0x80000000u;
            std::vector<uint32_t> plan;
            plan.reserve((size_t)s.n_loci + vars.size());
            std::vector<OutRec> overrides;   // reference records replaced by a reallocated MNV's (rare): planned as variants appended to `vars`
            auto plan_var = [&](size_t k) { if (mine(vars[k].r.position)) plan.push_back(kVar | (uint32_t)k); };
            const size_t n_vars = vars.size();
            for (int64_t i = 0; i < s.n_loci; i++) {
                const int32_t pos = s.has_positions ? s.h_positions[(size_t)i] : s.first_position + (int32_t)i;
                if (pos > cleared_to) break;
                while (vi < n_vars && vars[vi].r.position < pos) plan_var(vi++);
                bool variant_here = false;
                while (vi < n_vars && vars[vi].r.position == pos) { plan_var(vi++); variant_here = true; }
                if (!refs.empty() && valid[(size_t)i] && !variant_here && mine(pos)) {
                    auto ov = ref_override.empty() ? ref_override.end() : ref_override.find(pos);
                    if (ov == ref_override.end()) plan.push_back((uint32_t)i);
                    else { plan.push_back(kVar | (uint32_t)(n_vars + overrides.size())); overrides.push_back(ov->second); }
                }
            }
            while (vi < n_vars && vars[vi].r.position <= cleared_to) plan_var(vi++);
            const size_t base = h->h_out.size();
            h->h_out.resize(base + plan.size());
            h->h_out_ext.resize(base + plan.size());
            pb2_call_record* out_r = h->h_out.data() + base;
            pb2_call_record_ext* out_e = h->h_out_ext.data() + base;
            auto fill = [&](size_t a, size_t b) {
                for (size_t j = a; j < b; j++) {
                    const uint32_t e = plan[j];
                    if (e & kVar) {
                        const size_t k = e & ~kVar;
                        const OutRec& o = k < n_vars ? vars[k] : overrides[k - n_vars];
                        out_r[j] = o.r; out_e[j] = o.e;
                    } else if (collapsed.empty()) {
                        out_r[j] = refs[e]; out_e[j] = zero_ext;
                    } else {
                        const OutRec o = with_totals(OutRec{refs[e], zero_ext});
                        out_r[j] = o.r; out_e[j] = o.e;
                    }
                }
            };
            const size_t n_threads = plan.size() < (1u << 16) ? 1 : std::min<size_t>(8, std::max<unsigned>(1, std::thread::hardware_concurrency() / 4));
            if (n_threads <= 1) fill(0, plan.size());
            else {
                std::vector<std::thread> pool;
                for (size_t t = 0; t < n_threads; t++) pool.emplace_back(fill, plan.size() * t / n_threads, plan.size() * (t + 1) / n_threads);
                for (auto& t : pool) t.join();
            }
        }
    }
    {   // explicit alleles at positions no segment stages (e.g. an insertion before the first covered base)
        std::vector<OutRec> all;
        bool any_rest = false;
        for (size_t k = 0; k < explicit_called.size(); k++) any_rest |= !explicit_used[k];
        if (any_rest) {
            for (size_t k = 0; k < h->h_out.size(); k++) all.push_back(OutRec{h->h_out[k], h->h_out_ext[k]});
            for (size_t k = 0; k < explicit_called.size(); k++)
                if (!explicit_used[k]) {
                    OutRec o{explicit_called[k], explicit_ext[k]};
                    if (o.e.collapsed_total[0] == kCollapsedTotalTwice) o.e.collapsed_total[0] = 0;   // no staged position: no totals to double
                    all.push_back(o);
                }
            std::stable_sort(all.begin(), all.end(), [&](const OutRec& a, const OutRec& b) {
                return a.r.position != b.r.position ? a.r.position < b.r.position : (h->forced.empty() ? false : record_less_arena(a.r, b.r, h->arena));
            });
            h->h_out.clear(); h->h_out_ext.clear();
            for (auto& o : all) emit(o);
        }
    }
    germline_locus_pass(h);
    tr.mark("merge");
    // drop what this flush consumed: temporary segments, reads that end inside the cleared positions, dead candidates, used gapped counts
    for (size_t i = 0; i < h->segs.size();) {
        if (h->segs[i].temporary) { free_segment(h, h->segs[i]); h->segs.erase(h->segs.begin() + (long)i); } else i++;
    }
    h->cands.erase(std::remove_if(h->cands.begin(), h->cands.end(), [](const HostCand& c) { return !c.alive; }), h->cands.end());
    explicit_reindex(h);
    if (up_to_position < 0 && keep_reads) { h->cleared_through = 0; h->gapped_ref.clear(); h->last_trigger_key = 0; h->snv_explicit_ranges.clear(); }
    else if (up_to_position < 0) { clear_reads(h); h->cleared_through = 0; h->gapped_ref.clear(); h->triggers.clear(); h->last_trigger_key = 0; h->push_last_key = 0; h->snv_explicit_ranges.clear(); rearm_forced(h); }
    else if (reads_path && cleared_to > h->cleared_through) {
        const int rc = compact_reads(h, cleared_to);
        if (rc != PB2_OK) return rc;
        h->cleared_through = cleared_to;
        for (auto it = h->gapped_ref.begin(); it != h->gapped_ref.end();) { if (it->first <= cleared_to) it = h->gapped_ref.erase(it); else ++it; }
        h->snv_explicit_ranges.erase(std::remove_if(h->snv_explicit_ranges.begin(), h->snv_explicit_ranges.end(),
                                                    [&](const std::pair<int32_t, int32_t>& r) { return r.second <= cleared_to; }), h->snv_explicit_ranges.end());
    }
    tr.mark("cleanup");
    *out = h->h_out.data();
    *n = (int64_t)h->h_out.size();
    return PB2_OK;
}

extern "C" int pb2_flush_ext(pb2_handle* h, const pb2_call_record_ext** out, int64_t* n) {
    if (!h || !out || !n) return fail(h, PB2_ERR_ARG, "pb2_flush_ext: null argument");
    *out = h->h_out_ext.data();
    *n = (int64_t)h->h_out_ext.size();
    return PB2_OK;
}

extern "C" int pb2_set_forced_alleles(pb2_handle* h, const pb2_candidate* alleles, int32_t n, const uint8_t* arena, int64_t arena_len) {
    if (!h || n < 0 || (n > 0 && (!alleles || !arena))) return fail(h, PB2_ERR_ARG, "pb2_set_forced_alleles: bad argument");
    std::set<std::tuple<int32_t, std::string, std::string>> forced;
    std::vector<std::tuple<int32_t, std::string, std::string>> order;
    for (int32_t i = 0; i < n; i++) {
        const pb2_candidate& c = alleles[i];
        if (c.position < 1 || c.ref_len == 0 || c.alt_len == 0 || (int64_t)c.allele_offset + c.ref_len + c.alt_len > arena_len)
            return fail(h, PB2_ERR_ARG, "pb2_set_forced_alleles: allele " + std::to_string(i) + " is malformed");
        const char* b = reinterpret_cast<const char*>(arena) + c.allele_offset;
        std::string ref(b, c.ref_len), alt(b + c.ref_len, c.alt_len);
        if (ref.size() == alt.size() && ref.size() > 1 && !h->cfg.call_mnvs)
            return fail(h, PB2_ERR_UNSUPPORTED, "pb2_set_forced_alleles: a forced MNV needs call_mnvs=1 (its reallocation targets are count-based SNVs otherwise)");
        if (forced.insert(std::make_tuple(c.position, ref, alt)).second) order.push_back(std::make_tuple(c.position, ref, alt));
    }
    h->forced_order = std::move(order);
    rearm_forced(h);
    explicit_release_resident(h);
    release_resident_graph(h);
    return PB2_OK;
}
static void rearm_forced(pb2_handle* h) {
    h->forced_pending.clear();
    h->forced_positions.clear();
    h->forced.clear();   // Factory.SelectForcedAllele (Factory.cs:270-285): only the alleles inside the intervals are forced at all
    for (auto& f : h->forced_order) {
        const int32_t pos = std::get<0>(f);
        if (h->have_intervals) {   // _intervalSet.ContainsPosition (SmallVariantCaller.cs:58-61)
            bool in = false;
            for (size_t i = 0; i < h->iv_start.size() && !in; i++) in = pos >= h->iv_start[i] && pos <= h->iv_end[i];
            if (!in) continue;
        }
        h->forced.insert(f);
        h->forced_pending[pos].push_back({std::get<1>(f), std::get<2>(f)});
        h->forced_positions.push_back(pos);
    }
}

extern "C" int pb2_push_candidates(pb2_handle* h, const pb2_candidate* cands, int32_t n, const uint8_t* arena, int64_t arena_len) {
    if (!h || n < 0 || (n > 0 && (!cands || !arena))) return fail(h, PB2_ERR_ARG, "pb2_push_candidates: bad argument");
    for (int32_t i = 0; i < n; i++) {
        const pb2_candidate& c = cands[i];
        if (c.position <= 0) return fail(h, PB2_ERR_ARG, "Coordinate is invalid.");                        // CandidateAllele.cs:96-110
        if (c.ref_len == 0) return fail(h, PB2_ERR_ARG, "Reference is empty.");
        if (c.alt_len == 0) return fail(h, PB2_ERR_ARG, "Alternate is empty.");
        if (c.type > CAT_MNV) return fail(h, PB2_ERR_ARG, "reference candidates are not tracked");           // RegionState.cs:96-97
        if ((int64_t)c.allele_offset + c.ref_len + c.alt_len > arena_len) return fail(h, PB2_ERR_ARG, "pb2_push_candidates: allele outside the arena");
        if (c.type == CAT_SNV && !h->cfg.call_mnvs)
            return fail(h, PB2_ERR_ARG, "pb2_push_candidates: with CallMNVs off, SNV support is taken from the pileup counts; push SNV candidates only with call_mnvs=1");
        HostCand hc;
        hc.position = c.position; hc.type = c.type;
        hc.open_left = (c.open_flags & 1) != 0; hc.open_right = (c.open_flags & 2) != 0;
        hc.ref.assign(reinterpret_cast<const char*>(arena) + c.allele_offset, c.ref_len);
        hc.alt.assign(reinterpret_cast<const char*>(arena) + c.allele_offset + c.ref_len, c.alt_len);
        for (int k = 0; k < 3; k++) { hc.support[k] = c.support[k]; hc.well_anchored[k] = c.well_anchored[k]; }
        for (int k = 0; k < 8; k++) hc.collapsed_mut[k] = c.collapsed_mut[k];
        explicit_add_candidate(h, hc);
    }
    explicit_release_resident(h);
    release_resident_graph(h);
    return PB2_OK;
}

extern "C" int pb2_allele_arena(pb2_handle* h, const uint8_t** arena, int64_t* len) {
    if (!h || !arena || !len) return fail(h, PB2_ERR_ARG, "pb2_allele_arena: null argument");
    *arena = h->arena.data();
    *len = (int64_t)h->arena.size();
    return PB2_OK;
}

extern "C" int pb2_get_counts(pb2_handle* h, int32_t position0, int32_t n, int32_t* out) {
    if (!h || !out || n < 0) return fail(h, PB2_ERR_ARG, "pb2_get_counts: bad argument");
    CU(h, cudaSetDevice(h->device));
    memset(out, 0, sizeof(int32_t) * (size_t)n * kNumBins);   // positions nobody staged read as 0 (RegionStateManager.cs:224-225)
    {
        // staged reads are expanded over the requested window only (no intervals filter: counts exist for every position)
        const bool iv = h->have_intervals;
        h->have_intervals = false;
        const int rc = stage_reads_segment(h, position0 + n - 1, position0);
        h->have_intervals = iv;
        if (rc != PB2_OK) return rc;
    }
    for (auto& s : h->segs) {
        if (s.from_reads) continue;   // the same reads were just staged over the requested window
        int32_t* d_counts = nullptr;
        CU(h, cudaMalloc(&d_counts, sizeof(int32_t) * (size_t)s.n_loci * kNumBins));
        if (s.pv_data != nullptr) {   // PVERT: the 198-bin gather over every locus of the window
            int32_t* d_req = nullptr;
            std::vector<int32_t> req((size_t)s.n_loci);
            for (int64_t i = 0; i < s.n_loci; i++) req[(size_t)i] = (int32_t)i;
            CU(h, cudaMalloc(&d_req, sizeof(int32_t) * req.size()));
            CU(h, cudaMemcpyAsync(d_req, req.data(), sizeof(int32_t) * req.size(), cudaMemcpyHostToDevice, h->stream));
            cudaError_t e = launch_pvert_gather(pvert_view(s), d_req, (int32_t)s.n_loci, d_counts, nullptr, h->dcfg.min_bq, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            cudaFree(d_req);
            if (e != cudaSuccess) { cudaFree(d_counts); h->error = std::string("pvert_gather: ") + cudaGetErrorString(e); return PB2_ERR_CUDA; }
        } else {
            const bool was_called = s.called;
            const unsigned long long vc = s.h_var_count, ec = s.h_exc_count;
            int rc = run_segment(h, s, d_counts, nullptr);
            s.called = was_called; s.h_var_count = vc; s.h_exc_count = ec;
            if (rc != PB2_OK) { cudaFree(d_counts); return rc; }
        }
        std::vector<int32_t> hc((size_t)s.n_loci * kNumBins);
        CU(h, cudaMemcpy(hc.data(), d_counts, sizeof(int32_t) * hc.size(), cudaMemcpyDeviceToHost));
        cudaFree(d_counts);
        for (int64_t i = 0; i < s.n_loci; i++) {
            const int32_t pos = s.has_positions ? s.h_positions[(size_t)i] : s.first_position + (int32_t)i;
            const int64_t k = (int64_t)pos - position0;
            if (k >= 0 && k < n) memcpy(out + k * kNumBins, hc.data() + i * kNumBins, sizeof(int32_t) * kNumBins);
        }
    }
    for (size_t i = 0; i < h->segs.size();) {
        if (h->segs[i].temporary) { free_segment(h, h->segs[i]); h->segs.erase(h->segs.begin() + (long)i); } else i++;
    }
    return PB2_OK;
}

extern "C" int pb2_stage_stats(pb2_handle* h, int64_t* staged_bytes, int64_t* rows, double* stage_ms) {
    if (!h) return PB2_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    if (staged_bytes) *staged_bytes = h->last_stage_bytes;
    if (rows) *rows = h->last_stage_rows;
    if (stage_ms) {
        *stage_ms = 0;
        if (h->have_stage_events) {
            float ms = 0;
            CU(h, cudaEventSynchronize(h->ev_stage1));
            CU(h, cudaEventElapsedTime(&ms, h->ev_stage0, h->ev_stage1));
            *stage_ms = ms;
        }
    }
    return PB2_OK;
}

extern "C" int pb2_stats(pb2_handle* h, int64_t* hot_kernel_launches, double* hot_kernel_ms, int64_t* total_kernel_launches) {
    if (!h) return PB2_ERR_ARG;
    if (hot_kernel_launches) *hot_kernel_launches = h->hot_launches;
    if (hot_kernel_ms) *hot_kernel_ms = h->hot_ms;
    if (total_kernel_launches) *total_kernel_launches = h->total_launches;
    h->hot_launches = 0; h->hot_ms = 0; h->total_launches = 0;
    return PB2_OK;
}
