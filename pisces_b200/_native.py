"""ctypes binding of libpisces_b200.so (C ABI in include/pisces_b200.h). Fails loudly if the library is missing: there is no
Python or CPU implementation of the path behind it."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpisces_b200.so")


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("min_base_call_quality", C.c_int32), ("min_frequency", C.c_float), ("min_frequency_filter", C.c_float),
        ("target_lod_frequency", C.c_float), ("max_variant_qscore", C.c_int32), ("min_variant_qscore", C.c_int32),
        ("variant_qscore_filter", C.c_int32), ("max_genotype_qscore", C.c_int32), ("min_genotype_qscore", C.c_int32),
        ("low_genotype_quality_filter", C.c_int32), ("min_coverage", C.c_int32), ("low_depth_filter", C.c_int32),
        ("rmxn_max_repeat_len", C.c_int32), ("rmxn_min_repetitions", C.c_int32), ("rmxn_frequency_limit", C.c_float),
        ("forced_noise_level", C.c_int32), ("noise_model", C.c_int32), ("strand_bias_acceptance", C.c_float),
        ("strand_bias_model", C.c_int32), ("filter_single_strand", C.c_int32), ("no_call_filter", C.c_float), ("ploidy", C.c_int32),
        ("tracked_anchor_size", C.c_int32), ("output_gvcf", C.c_int32), ("expect_stitched", C.c_int32), ("expect_collapsed", C.c_int32),
        ("want_sum_base_quality", C.c_int32), ("collapse", C.c_int32), ("call_mnvs", C.c_int32), ("indel_repeat_filter", C.c_int32),
        ("max_size_mnv", C.c_int32), ("max_gap_mnv", C.c_int32), ("collapse_freq_threshold", C.c_float),
        ("collapse_freq_ratio_threshold", C.c_float), ("exclude_mnvs_from_collapsing", C.c_int32), ("skip_validation", C.c_int32),
        ("diploid_minor_vf", C.c_float), ("diploid_major_vf", C.c_float), ("diploid_sum_vf_multiallelic", C.c_float), ("is_male", C.c_int32),
        ("amplicon_bias_filter", C.c_float), ("reserved", C.c_int32 * 2)]


class PileupCsr(C.Structure):
    _fields_ = [("n_loci", C.c_int64), ("first_position", C.c_int32), ("positions", C.c_void_p), ("offsets", C.c_void_p),
                ("code", C.c_void_p), ("qual", C.c_void_p), ("anchor", C.c_void_p), ("ref_bases", C.c_void_p),
                ("layout", C.c_int32), ("reserved", C.c_int32), ("n_flags", C.c_int64), ("flag_index", C.c_void_p), ("flag_bits", C.c_void_p)]


class ReadBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("pos0", C.c_void_p), ("flag", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar", C.c_void_p),
                ("seq_off", C.c_void_p), ("bases", C.c_void_p), ("quals", C.c_void_p), ("base_dirs", C.c_void_p), ("collapsed", C.c_void_p),
                ("amplicon", C.c_void_p)]


class PackedReadBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("pos0", C.c_void_p), ("flag", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar", C.c_void_p), ("seq_off", C.c_void_p),
                ("seq", C.c_void_p), ("n_exceptions", C.c_int64), ("exc_index", C.c_void_p), ("exc_base", C.c_void_p), ("exc_qual", C.c_void_p),
                ("base_dirs", C.c_void_p), ("collapsed", C.c_void_p), ("amplicon", C.c_void_p), ("cigar_ops", C.c_void_p), ("n_cigar_total", C.c_int64),
                ("n_seq_total", C.c_int64)]


class Shard(C.Structure):
    _fields_ = [("own_lo", C.c_int32), ("own_hi", C.c_int32), ("stage_lo", C.c_int32), ("stage_hi", C.c_int32), ("read_first", C.c_int64), ("read_end", C.c_int64)]


class Candidate(C.Structure):
    _fields_ = [("position", C.c_int32), ("type", C.c_uint8), ("open_flags", C.c_uint8), ("ref_len", C.c_uint16), ("alt_len", C.c_uint16),
                ("reserved", C.c_uint16), ("allele_offset", C.c_uint32), ("support", C.c_int32 * 3), ("well_anchored", C.c_int32 * 3),
                ("collapsed_mut", C.c_int32 * 8)]


assert C.sizeof(Candidate) == 72
CANDIDATE_DTYPE = [("position", "<i4"), ("type", "u1"), ("open_flags", "u1"), ("ref_len", "<u2"), ("alt_len", "<u2"), ("reserved", "<u2"),
                   ("allele_offset", "<u4"), ("support", "<i4", (3,)), ("well_anchored", "<i4", (3,)), ("collapsed_mut", "<i4", (8,))]


class CallRecord(C.Structure):
    _fields_ = [("position", C.c_int32), ("type", C.c_uint8), ("genotype", C.c_uint8), ("sb_flags", C.c_uint8), ("open_flags", C.c_uint8),
                ("filters", C.c_uint16), ("noise_level", C.c_uint16), ("variant_qscore", C.c_int32), ("genotype_qscore", C.c_int32),
                ("total_coverage", C.c_int32), ("coverage_by_direction", C.c_int32 * 3), ("support_by_direction", C.c_int32 * 3),
                ("allele_support", C.c_int32), ("reference_support", C.c_int32), ("num_no_calls", C.c_int32),
                ("fraction_no_calls", C.c_float), ("allele_bytes", C.c_uint32), ("ref_len", C.c_uint16), ("alt_len", C.c_uint16),
                ("sum_base_quality", C.c_double), ("bias_score", C.c_double), ("gatk_bias_score", C.c_double)]


assert C.sizeof(CallRecord) == 96

# numpy view of pb2_call_record
RECORD_DTYPE = [("position", "<i4"), ("type", "u1"), ("genotype", "u1"), ("sb_flags", "u1"), ("open_flags", "u1"), ("filters", "<u2"),
                ("noise_level", "<u2"), ("variant_qscore", "<i4"), ("genotype_qscore", "<i4"), ("total_coverage", "<i4"),
                ("coverage_by_direction", "<i4", (3,)), ("support_by_direction", "<i4", (3,)), ("allele_support", "<i4"),
                ("reference_support", "<i4"), ("num_no_calls", "<i4"), ("fraction_no_calls", "<f4"), ("allele_bytes", "<u4"),
                ("ref_len", "<u2"), ("alt_len", "<u2"), ("sum_base_quality", "<f8"), ("bias_score", "<f8"), ("gatk_bias_score", "<f8")]

RECORD_EXT_DTYPE = [("collapsed_mut", "<i4", (8,)), ("collapsed_total", "<i4", (8,)), ("well_anchored_support", "<i4", (3,)), ("phase_set_index", "<i4")]

EXPORTS = ["pb2_default_config", "pb2_create", "pb2_destroy", "pb2_last_error", "pb2_device_count", "pb2_set_reference", "pb2_set_intervals",
           "pb2_push_pileup", "pb2_push_pileup_device", "pb2_push_reads", "pb2_push_reads_packed", "pb2_pack_reads", "pb2_pack_pileup", "pb2_stage_reads", "pb2_push_candidates", "pb2_set_forced_alleles", "pb2_allele_arena", "pb2_call_resident", "pb2_call_resident_async", "pb2_resident_sync", "pb2_set_resident_sink", "pb2_sink_sort", "pb2_resident_results", "pb2_flush", "pb2_flush_resident", "pb2_flush_ext",
           "pb2_get_counts", "pb2_reset", "pb2_shard_plan", "pb2_set_owned_range", "pb2_stats", "pb2_stage_stats", "pb2_stream", "pb2_totals", "pb2_vcf_format",
           "pb2_bam_open", "pb2_bam_close", "pb2_bam_last_error", "pb2_bam_header", "pb2_bam_next_batch", "pb2_bam_next_batch_packed", "pb2_bam_batch_amplicons",
           "pb2_bam_amplicon_names"]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C pisces_b200/csrc). pisces_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.pb2_default_config.argtypes = [C.POINTER(Config)]
    L.pb2_default_config.restype = None
    L.pb2_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    L.pb2_destroy.argtypes = [H]
    L.pb2_destroy.restype = None
    L.pb2_last_error.argtypes = [H]
    L.pb2_last_error.restype = C.c_char_p
    L.pb2_set_reference.argtypes = [H, C.c_char_p, C.c_char_p, C.c_int64]
    L.pb2_set_intervals.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int32]
    L.pb2_push_pileup.argtypes = [H, C.POINTER(PileupCsr)]
    L.pb2_push_pileup_device.argtypes = [H, C.POINTER(PileupCsr)]
    L.pb2_push_reads.argtypes = [H, C.POINTER(ReadBatch)]
    L.pb2_push_reads_packed.argtypes = [H, C.POINTER(PackedReadBatch)]
    L.pb2_pack_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.pb2_pack_reads.restype = C.c_int64
    L.pb2_pack_pileup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.pb2_pack_pileup.restype = C.c_int64
    L.pb2_stage_reads.argtypes = [H]
    L.pb2_push_candidates.argtypes = [H, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
    L.pb2_set_forced_alleles.argtypes = [H, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
    L.pb2_allele_arena.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.pb2_call_resident.argtypes = [H, C.POINTER(C.c_int64)]
    L.pb2_call_resident_async.argtypes = [H]
    L.pb2_resident_sync.argtypes = [H, C.POINTER(C.c_int64)]
    L.pb2_set_resident_sink.argtypes = [H, C.c_void_p, C.c_int64, C.c_int32]
    L.pb2_sink_sort.argtypes = [H]
    L.pb2_resident_results.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.pb2_flush.argtypes = [H, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.pb2_flush_resident.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.pb2_flush_ext.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.pb2_get_counts.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p]
    L.pb2_reset.argtypes = [H]
    L.pb2_stats.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.pb2_shard_plan.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Shard)]
    L.pb2_set_owned_range.argtypes = [H, C.c_int32, C.c_int32]
    L.pb2_stage_stats.argtypes = [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.pb2_vcf_format.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_int64)]
    L.pb2_bam_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.pb2_bam_close.argtypes = [C.c_void_p]
    L.pb2_bam_close.restype = None
    L.pb2_bam_last_error.argtypes = [C.c_void_p]
    L.pb2_bam_last_error.restype = C.c_char_p
    L.pb2_bam_header.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.POINTER(C.c_char_p)), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32)]
    L.pb2_bam_next_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(ReadBatch), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.pb2_bam_next_batch_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(PackedReadBatch), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.pb2_bam_batch_amplicons.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32)]
    L.pb2_bam_amplicon_names.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.POINTER(C.c_char_p))]
    L.pb2_totals.argtypes = [H, C.POINTER(C.c_int64)]
    L.pb2_stream.argtypes = [H]
    L.pb2_stream.restype = C.c_void_p
    _lib = L
    return L
