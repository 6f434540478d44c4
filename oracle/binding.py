"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/_build/liboracle.so (built by oracle/Makefile)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
            for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp", ".h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("min_base_call_quality", "min_map_quality", "remove_duplicates", "only_proper_pairs")] + \
        [(n, C.c_float) for n in ("min_frequency", "min_frequency_filter", "target_lod_frequency")] + \
        [(n, C.c_int32) for n in ("max_vq", "min_vq", "vq_filter", "max_gq", "min_gq", "low_gq_filter", "min_coverage", "low_depth_filter",
                                  "indel_repeat_filter", "rmxn_max_repeat_len", "rmxn_min_repetitions")] + \
        [("rmxn_freq_limit", C.c_float)] + \
        [(n, C.c_int32) for n in ("ploidy", "forced_noise_level", "noise_model")] + \
        [("sb_acceptance", C.c_float)] + \
        [(n, C.c_int32) for n in ("sb_model", "filter_single_strand")] + \
        [("no_call_filter", C.c_float)] + \
        [(n, C.c_int32) for n in ("call_mnvs", "max_size_mnv", "max_gap_mnv", "collapse")] + \
        [(n, C.c_float) for n in ("collapse_freq_threshold", "collapse_freq_ratio_threshold")] + \
        [(n, C.c_int32) for n in ("exclude_mnvs_from_collapsing", "tracked_anchor_size", "output_gvcf", "source_is_stitched", "source_is_collapsed", "apply_validation")] + \
        [(n, C.c_float) for n in ("diploid_minor_vf", "diploid_major_vf", "diploid_sum_vf_multiallelic")] + [("is_male", C.c_int32), ("amplicon_bias_filter", C.c_float)]


class ReadStruct(C.Structure):
    _fields_ = [("pos0", C.c_int32), ("flag", C.c_int32), ("mapq", C.c_int32), ("n_cigar", C.c_int32), ("cigar", C.POINTER(C.c_uint32)),
                ("l_seq", C.c_int32), ("seq", C.c_char_p), ("qual", C.POINTER(C.c_uint8)), ("has_tags", C.c_int32), ("xd", C.c_char_p),
                ("xr", C.c_char_p), ("has_xv", C.c_int32), ("xv", C.c_int32), ("has_xw", C.c_int32), ("xw", C.c_int32)]


class Record(C.Structure):
    _fields_ = [("pos", C.c_int32), ("type", C.c_int32), ("genotype", C.c_int32), ("gq", C.c_int32), ("vq", C.c_int32),
                ("filter_mask", C.c_uint32), ("n_filters", C.c_int32), ("filters", C.c_int32 * 8),
                ("noise_level", C.c_int32), ("total_coverage", C.c_int32), ("sum_base_quality", C.c_double),
                ("cov", C.c_int32 * 3), ("support", C.c_int32 * 3), ("well_anchored", C.c_int32 * 3),
                ("allele_support", C.c_int32), ("ref_support", C.c_int32), ("num_no_calls", C.c_int32),
                ("fraction_no_calls", C.c_float), ("frequency", C.c_float), ("bias_score", C.c_double), ("gatk_bias_score", C.c_double),
                ("bias_acceptable", C.c_int32), ("var_both_strands", C.c_int32), ("cov_both_strands", C.c_int32), ("forced", C.c_int32),
                ("collapsed_mut", C.c_int32 * 8), ("collapsed_total", C.c_int32 * 8), ("ref_len", C.c_int32), ("alt_len", C.c_int32),
                ("has_amplicon_bias", C.c_int32), ("amplicon_bias_detected", C.c_int32),
                ("n_amp_support", C.c_int32), ("amp_support_names", C.c_int32 * 6), ("amp_support_counts", C.c_int32 * 6),
                ("n_amp_coverage", C.c_int32), ("amp_coverage_names", C.c_int32 * 6), ("amp_coverage_counts", C.c_int32 * 6)]


# enums (src/lib/Pisces.Domain/Types/*.cs)
A, G, Cc, T, N, DEL = 0, 1, 2, 3, 4, 5
FWD, REV, STITCHED = 0, 1, 2
SNV, INSERTION, DELETION, MNV, REFERENCE = 0, 1, 2, 3, 4
FILTERS = ["StrandBias", "PoolBias", "AmpliconBias", "LowVariantQscore", "LowDepth", "LowVariantFrequency", "LowGenotypeQuality",
           "IndelRepeatLength", "MultiAllelicSite", "RMxN", "ForcedReport", "OffTarget", "NoCall", "Unknown"]
GENOTYPES = ["HeterozygousAlt1Alt2", "Alt12LikeNoCall", "HeterozygousAltRef", "HomozygousAlt", "HomozygousRef", "RefLikeNoCall",
             "AltLikeNoCall", "RefAndNoCall", "AltAndNoCall", "HemizygousRef", "HemizygousAlt", "HemizygousNoCall", "Others"]
_CIGAR_OPS = "MIDNSHP=X"

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.po_last_error.restype = C.c_char_p
        L.po_caller_create.restype = C.c_void_p
        L.po_caller_create.argtypes = [C.POINTER(Config), C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32]
        L.po_caller_destroy.argtypes = [C.c_void_p]
        L.po_caller_add_forced.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_char_p]
        for f in ("po_caller_add_read", "po_caller_add_read_counts_only", "po_caller_add_read_candidates_only"):
            getattr(L, f).argtypes = [C.c_void_p, C.POINTER(ReadStruct)]
        L.po_caller_add_reads_soa.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 9
        L.po_caller_add_reads_soa_amplicons.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10
        L.po_caller_add_reads_soa_counts_only.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 9
        L.po_caller_add_pileup.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.po_caller_add_candidate.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                              C.POINTER(C.c_int32)]
        L.po_caller_finish.argtypes = [C.c_void_p]
        L.po_caller_num_records.argtypes = [C.c_void_p]
        L.po_caller_get_record.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Record)]
        for f in ("po_caller_record_ref", "po_caller_record_alt"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int32]
            getattr(L, f).restype = C.c_char_p
        L.po_caller_num_write_batches.argtypes = [C.c_void_p]
        L.po_caller_write_batch.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.po_caller_total_called.argtypes = [C.c_void_p]
        L.po_caller_total_collapsed.argtypes = [C.c_void_p]
        L.po_get_allele_count.argtypes = [C.c_void_p] + [C.c_int32] * 7
        L.po_get_sum_base_quality.argtypes = [C.c_void_p] + [C.c_int32] * 6
        L.po_get_sum_base_quality.restype = C.c_double
        L.po_get_collapsed_count.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.po_set_allele_count.argtypes = [C.c_void_p] + [C.c_int32] * 5
        L.po_add_gapped_ref_count.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.po_dump_counts.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        L.po_num_candidates_at.argtypes = [C.c_void_p, C.c_int32]
        L.po_get_candidate_at.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(C.c_int32)]
        L.po_process_allele.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(Record)]
        L.po_raw_vq.argtypes = [C.c_int32] * 3
        L.po_raw_vq.restype = C.c_double
        L.po_vq.argtypes = [C.c_int32] * 4
        L.po_pvalue.argtypes = [C.c_int32] * 3
        L.po_pvalue.restype = C.c_double
        L.po_exact_spanning_read_direction.argtypes = [C.c_int32] * 5 + [C.c_char_p, C.c_char_p]
        L.po_exact_spanning_read_direction.restype = C.c_int32
        L.po_amplicon_bias.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.c_void_p]
        L.po_amplicon_bias.restype = C.c_int32
        L.po_poisson_cdf.argtypes = [C.c_double, C.c_double]
        L.po_poisson_cdf.restype = C.c_double
        L.po_mathnet_gamma_lower_regularized.argtypes = [C.c_double, C.c_double]
        L.po_mathnet_gamma_lower_regularized.restype = C.c_double
        L.po_mathnet_gamma_ln.argtypes = [C.c_double]
        L.po_mathnet_gamma_ln.restype = C.c_double
        L.po_strand_bias.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_double, C.c_double, C.c_int32, C.POINTER(C.c_double)]
        L.po_somatic_gq.argtypes = [C.c_int32] * 5 + [C.c_float, C.c_int32, C.c_int32]
        L.po_somatic_genotype.argtypes = [C.c_int32] * 4 + [C.c_float, C.c_int32]
        L.po_anchor_adjusted_count.argtypes = [C.POINTER(C.c_int32)] + [C.c_int32] * 5
        L.po_mathnet_binomial_cdf.argtypes = [C.c_double, C.c_int32, C.c_double]
        L.po_mathnet_binomial_cdf.restype = C.c_double
        L.po_mathnet_binomial_probability_ln.argtypes = [C.c_double, C.c_int32, C.c_int32]
        L.po_mathnet_binomial_probability_ln.restype = C.c_double
        L.po_diploid_gq.argtypes = [C.c_int32] * 5
        L.po_haploid_gq.argtypes = [C.c_int32] * 5
        L.po_genotype_locus.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                        C.POINTER(C.c_char_p), C.c_int32, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32)]
        L.po_ploidy_for_chr.argtypes = [C.c_int32, C.c_int32, C.c_char_p]
        _lib = L
    return _lib


def genotype_locus(ploidy, alleles, min_depth=100, minor_vf=0.20, major_vf=0.70, sum_vf=0.80):
    """alleles: list of (type, allele_support, total_coverage, ref_support[, "REF>ALT"]). Returns (genotype name, pruned flags, multiallelic flag, gqs)."""
    n = len(alleles)
    arr = lambda k: (C.c_int32 * max(n, 1))(*[a[k] for a in alleles])
    names = (C.c_char_p * max(n, 1))(*[(a[4].encode() if len(a) > 4 else b"A>C") for a in alleles])
    pruned, multi, gq = (C.c_int32 * max(n, 1))(), C.c_int32(), (C.c_int32 * max(n, 1))()
    g = lib().po_genotype_locus(ploidy, n, arr(0), arr(1), arr(2), arr(3), names, min_depth, minor_vf, major_vf, sum_vf, pruned, C.byref(multi), gq)
    return GENOTYPES[g], list(pruned)[:n], multi.value, list(gq)[:n]


def default_config(**kw):
    c = Config()
    lib().po_default_config(C.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def parse_cigar(s):
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | _CIGAR_OPS.index(ch))
            num = ""
    return out


class SimpleRead:
    """A read as the reference's tests build them (ReadTestHelper.CreateRead, TestUtilities/ReadTestHelper.cs:72)."""

    def __init__(self, pos, seq, cigar, quals=30, flag=0, mapq=10, xd=None, xr=None, xv=None, xw=None, has_tags=None):
        """pos is the 1-based Read.Position."""
        self.pos0 = pos - 1
        self.seq = seq
        self.cigar = parse_cigar(cigar) if isinstance(cigar, str) else list(cigar)
        self.quals = [quals] * len(seq) if isinstance(quals, int) else list(quals)
        self.flag, self.mapq, self.xd, self.xr, self.xv, self.xw = flag, mapq, xd, xr, xv, xw
        self.has_tags = has_tags if has_tags is not None else any(v is not None for v in (xd, xr, xv, xw))

    def to_struct(self):
        r = ReadStruct()
        r.pos0, r.flag, r.mapq = self.pos0, self.flag, self.mapq
        self._cig = (C.c_uint32 * max(1, len(self.cigar)))(*self.cigar)
        r.n_cigar, r.cigar = len(self.cigar), self._cig
        self._seq = self.seq.encode()
        r.l_seq, r.seq = len(self.seq), self._seq
        self._q = (C.c_uint8 * max(1, len(self.quals)))(*self.quals)
        r.qual = self._q
        r.has_tags = int(self.has_tags)
        r.xd = self.xd.encode() if self.xd is not None else None
        r.xr = self.xr.encode() if self.xr is not None else None
        r.has_xv, r.xv = int(self.xv is not None), self.xv or 0
        r.has_xw, r.xw = int(self.xw is not None), self.xw or 0
        return r


class Caller:
    def __init__(self, cfg=None, chr_name="chr1", seq="", intervals=None):
        self.L = lib()
        cfg = cfg or default_config()
        self.cfg = cfg
        seqb = seq.encode() if isinstance(seq, str) else bytes(seq)
        if intervals is None:
            h = self.L.po_caller_create(C.byref(cfg), chr_name.encode(), seqb, len(seqb), None, None, -1)
        else:
            s = (C.c_int32 * max(1, len(intervals)))(*[a for a, _ in intervals])
            e = (C.c_int32 * max(1, len(intervals)))(*[b for _, b in intervals])
            h = self.L.po_caller_create(C.byref(cfg), chr_name.encode(), seqb, len(seqb), s, e, len(intervals))
        if not h:
            raise RuntimeError(self.L.po_last_error().decode())
        self.h = C.c_void_p(h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.po_caller_destroy(self.h)
            self.h = None

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.po_last_error().decode())

    def add_read(self, r, mode="full"):
        st = r.to_struct()
        f = {"full": self.L.po_caller_add_read, "counts": self.L.po_caller_add_read_counts_only,
             "candidates": self.L.po_caller_add_read_candidates_only}[mode]
        self._chk(f(self.h, C.byref(st)))

    def add_reads_soa(self, pos0, flag, cigar_off, cigar, seq_off, bases, quals, collapsed=None, xd_runs=None, counts_only=False, amplicon=None):
        """A struct of arrays of reads (pb2_read_batch layout) through the reference's per-read loop. amplicon: per-read XN name ids, -1 = no tag."""
        import numpy as np
        arrs = [np.ascontiguousarray(pos0, dtype=np.int32), np.ascontiguousarray(flag, dtype=np.uint16), np.ascontiguousarray(cigar_off, dtype=np.int64),
                np.ascontiguousarray(cigar, dtype=np.uint32), np.ascontiguousarray(seq_off, dtype=np.int64), np.ascontiguousarray(bases, dtype=np.uint8),
                np.ascontiguousarray(quals, dtype=np.uint8)]
        coll = None if collapsed is None else np.ascontiguousarray(collapsed, dtype=np.uint8)
        xd = None if xd_runs is None else np.ascontiguousarray(xd_runs, dtype=np.int32)
        extra = []
        if amplicon is not None:
            assert not counts_only
            amp = np.ascontiguousarray(amplicon, dtype=np.int32)
            assert len(amp) == len(arrs[0])
            f, extra = self.L.po_caller_add_reads_soa_amplicons, [amp.ctypes.data]
        else:
            f = self.L.po_caller_add_reads_soa_counts_only if counts_only else self.L.po_caller_add_reads_soa
        self._chk(f(self.h, len(arrs[0]), *[a.ctypes.data for a in arrs], None if coll is None else coll.ctypes.data,
                    None if xd is None else xd.ctypes.data, *extra))

    def add_pileup(self, offsets, code, qual, anchor, first_position=1, call_every=1):
        """Locus-major entries (pb2_pileup_csr semantics) through the reference's per-base operations."""
        import numpy as np
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        code, qual, anchor = (np.ascontiguousarray(x, dtype=np.uint8) for x in (code, qual, anchor))
        self._chk(self.L.po_caller_add_pileup(self.h, len(offsets) - 1, first_position, offsets.ctypes.data, code.ctypes.data, qual.ctypes.data,
                                              anchor.ctypes.data, call_every))

    def add_candidate(self, type_, pos, ref, alt, support=(0, 0, 0), well_anchored=(0, 0, 0), open_left=False, open_right=False, collapsed_mut=None):
        cm = (C.c_int32 * 8)(*(collapsed_mut or [0] * 8))
        self._chk(self.L.po_caller_add_candidate(self.h, type_, pos, ref.encode(), alt.encode(), (C.c_int32 * 3)(*support), (C.c_int32 * 3)(*well_anchored),
                                                 int(open_left), int(open_right), cm))

    def records_array(self):
        """All records as a numpy structured array (fast path for large outputs)."""
        import numpy as np
        n = self.L.po_caller_num_records(self.h)
        arr = (Record * n)()
        for i in range(n):
            self._chk(self.L.po_caller_get_record(self.h, i, C.byref(arr[i])))
        return np.ctypeslib.as_array(arr) if n else np.zeros(0)

    def add_forced(self, pos, ref, alt):
        self.L.po_caller_add_forced(self.h, pos, ref.encode(), alt.encode())

    def finish(self):
        self._chk(self.L.po_caller_finish(self.h))

    def records(self):
        out = []
        for i in range(self.L.po_caller_num_records(self.h)):
            r = Record()
            self._chk(self.L.po_caller_get_record(self.h, i, C.byref(r)))
            r.ref = self.L.po_caller_record_ref(self.h, i).decode()
            r.alt = self.L.po_caller_record_alt(self.h, i).decode()
            out.append(r)
        return out

    def write_batches(self):
        out = []
        b, e = C.c_int32(), C.c_int32()
        for i in range(self.L.po_caller_num_write_batches(self.h)):
            self.L.po_caller_write_batch(self.h, i, C.byref(b), C.byref(e))
            out.append((b.value, e.value))
        return out

    def count(self, pos, allele, direction, min_anchor=0, max_anchor=None, from_end=False, symmetric=False):
        return self.L.po_get_allele_count(self.h, pos, allele, direction, min_anchor, -1 if max_anchor is None else max_anchor, int(from_end), int(symmetric))

    def qsum(self, pos, allele, direction, min_anchor=0, max_anchor=None, from_end=False):
        return self.L.po_get_sum_base_quality(self.h, pos, allele, direction, min_anchor, -1 if max_anchor is None else max_anchor, int(from_end))

    def collapsed_count(self, pos, t):
        return self.L.po_get_collapsed_count(self.h, pos, t)

    def set_count(self, pos, allele, direction, anchor, value):
        self.L.po_set_allele_count(self.h, pos, allele, direction, anchor, value)

    def add_gapped_ref_count(self, pos, count):
        self.L.po_add_gapped_ref_count(self.h, pos, count)

    def dump_counts(self, pos0, n):
        import numpy as np
        na = 2 * self.cfg.tracked_anchor_size + 1
        out = np.zeros((n, 6, 3, na), dtype=np.int32)
        self._chk(self.L.po_dump_counts(self.h, pos0, n, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def candidates_at(self, pos):
        out = []
        for i in range(self.L.po_num_candidates_at(self.h, pos)):
            t, ol, orr = C.c_int32(), C.c_int32(), C.c_int32()
            sup, wa, cm = (C.c_int32 * 3)(), (C.c_int32 * 3)(), (C.c_int32 * 8)()
            rb, ab = C.create_string_buffer(512), C.create_string_buffer(512)
            self._chk(self.L.po_get_candidate_at(self.h, pos, i, C.byref(t), sup, wa, C.byref(ol), C.byref(orr), rb, ab, 512, cm))
            out.append(dict(type=t.value, pos=pos, ref=rb.value.decode(), alt=ab.value.decode(), support=list(sup), well_anchored=list(wa),
                            open_left=bool(ol.value), open_right=bool(orr.value), collapsed_mut=list(cm)))
        return out

    def process_allele(self, type_, pos, ref, alt, support=(0, 0, 0), well_anchored=(0, 0, 0), coverage_only=False):
        r = Record()
        s = (C.c_int32 * 3)(*support)
        w = (C.c_int32 * 3)(*well_anchored)
        self._chk(self.L.po_process_allele(self.h, type_, pos, ref.encode(), alt.encode(), s, w, int(coverage_only), C.byref(r)))
        return r


def strand_bias(cov, sup, q_noise, min_vf=0.01, acceptance=0.5, model=1):
    out = (C.c_double * 29)()
    lib().po_strand_bias((C.c_int32 * 3)(*cov), (C.c_int32 * 3)(*sup), q_noise, min_vf, acceptance, model, out)
    names = ["overall", "fwd", "rev", "stitched"]
    res = dict(bias=out[0], gatk=out[1], acceptable=bool(out[2]), var_both=bool(out[3]), cov_both=bool(out[4]))
    for i, n in enumerate(names):
        o = out[5 + 6 * i: 11 + 6 * i]
        res[n] = dict(fn=o[0], fp=o[1], vg=o[2], coverage=o[3], frequency=o[4], support=o[5])
    return res


def amplicon_bias(support, coverage, acceptance=0.01, max_qscore=100):
    """AmpliconBiasCalculator.CalculateAmpliconBias on (names, counts) pairs; names are ints (-1 = null), support names None = null array.
    Returns None for a null result, else dict(bias_detected, artifact, per_amplicon=[dict(...)])."""
    import numpy as np
    sn, sc = support
    cn, cc = coverage
    ns = -1 if sn is None else len(sn)
    a = lambda v: np.ascontiguousarray(v if v is not None and len(v) else [0], dtype=np.int32)   # noqa: E731
    sn_, sc_, cn_, cc_ = a(sn), a(sc), a(cn), a(cc)
    per = np.zeros((max(len(cn), 1), 8), dtype=np.float64)
    bd, art = C.c_int32(0), C.c_int32(-1)
    n = lib().po_amplicon_bias(sn_.ctypes.data, sc_.ctypes.data, ns, cn_.ctypes.data, cc_.ctypes.data, len(cn), acceptance, max_qscore, C.byref(bd),
                               C.byref(art), per.ctypes.data)
    if n < 0:
        return None
    keys = ("name", "frequency", "coverage", "observed_support", "expected_support", "chance_its_real", "qscore", "bias_detected")
    return dict(bias_detected=bool(bd.value), artifact=art.value, per_amplicon=[dict(zip(keys, per[i])) for i in range(n)])


def exact_spanning_read_direction(allele_type, start, end, cigar, directions, position=10, allele_length=4):
    """ExactCoverageCalculator on one ReadCoverageSummary: 0 Forward / 1 Reverse / 2 Stitched, or None when the read does not contribute."""
    r = lib().po_exact_spanning_read_direction(allele_type, position, allele_length, start, end, cigar.encode(), directions.encode())
    if r < -1:
        raise ValueError(f"po_exact_spanning_read_direction: {r}")
    return None if r == -1 else r
