"""Host-side mirror of the reference's plug-in interfaces for the hot path, over the C ABI.

  GpuStateManager  ~ IStateManager : IAlleleSource   (src/lib/Pisces.Processing/Interfaces/IStateManager.cs:8-15,
                                                      src/lib/Pisces.Domain/Interfaces/IAlleleSource.cs:8-26)
  GpuAlleleCaller  ~ IAlleleCaller                   (src/exe/Pisces/Interfaces/IAlleleCaller.cs:8-13)
  CalledAllele     ~ CalledAllele                    (src/lib/Pisces.Domain/Models/Alleles/CalledAllele.cs)
  Read             ~ Read                            (src/lib/Pisces.Domain/Models/Read.cs), the fields the path consumes

Method names and argument meaning follow the reference so that tests read like the reference's own tests. All compute happens in
libpisces_b200.so on the GPU; this module only marshals.
"""
import ctypes as C
import os
import enum

import numpy as np

from . import _native as N


class PiscesB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pisces_b200 error {code}: {msg}")
        self.code = code


class AlleleType(enum.IntEnum):      # Types/AlleleType.cs:5-10
    A = 0
    G = 1
    C = 2
    T = 3
    N = 4
    Deletion = 5


class DirectionType(enum.IntEnum):   # Types/DirectionType.cs
    Forward = 0
    Reverse = 1
    Stitched = 2


class AlleleCategory(enum.IntEnum):  # Types/AlleleCategory.cs
    Snv = 0
    Insertion = 1
    Deletion = 2
    Mnv = 3
    Reference = 4


class FilterType(enum.IntEnum):      # Types/FilterType.cs
    StrandBias = 0
    PoolBias = 1
    AmpliconBias = 2
    LowVariantQscore = 3
    LowDepth = 4
    LowVariantFrequency = 5
    LowGenotypeQuality = 6
    IndelRepeatLength = 7
    MultiAllelicSite = 8
    RMxN = 9
    ForcedReport = 10
    OffTarget = 11
    NoCall = 12


class Genotype(enum.IntEnum):        # Types/Genotype.cs
    HeterozygousAlt1Alt2 = 0
    Alt12LikeNoCall = 1
    HeterozygousAltRef = 2
    HomozygousAlt = 3
    HomozygousRef = 4
    RefLikeNoCall = 5
    AltLikeNoCall = 6
    RefAndNoCall = 7
    AltAndNoCall = 8


VariantCallerConfig = N.Config


def make_config(**kw):
    """pb2_default_config() + overrides (field names of pb2_config)."""
    c = N.Config()
    N.load().pb2_default_config(C.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(f"pb2_config has no field {k}")
        setattr(c, k, v)
    return c


_CIGAR_OPS = "MIDNSHP=X"


class Read:
    """The fields of Pisces.Domain.Models.Read the path consumes. position is the 1-based Read.Position."""

    def __init__(self, position, sequence, cigar, qualities=30, flag=0, base_directions=None, collapsed=None):
        self.position = position
        self.sequence = sequence
        if isinstance(cigar, str):
            ops, num = [], ""
            for ch in cigar:
                if ch.isdigit():
                    num += ch
                else:
                    ops.append((int(num) << 4) | _CIGAR_OPS.index(ch))
                    num = ""
            cigar = ops
        self.cigar = list(cigar)
        self.qualities = [qualities] * len(sequence) if isinstance(qualities, int) else list(qualities)
        self.flag = flag
        self.base_directions = base_directions   # Read.SequencedBaseDirectionMap or None
        self.collapsed = collapsed               # summary byte (see pb2_read_batch.collapsed) or None


class CalledAllele:
    """View of one pb2_call_record with the reference's property names."""

    def __init__(self, rec, arena=b""):
        self._r = rec
        self.ReferencePosition = int(rec["position"])
        self.Type = AlleleCategory(int(rec["type"]))
        self.Genotype = Genotype(int(rec["genotype"]))
        self.GenotypeQscore = int(rec["genotype_qscore"])
        self.VariantQscore = int(rec["variant_qscore"])
        self.Filters = [FilterType(i) for i in range(13) if int(rec["filters"]) >> i & 1]
        self.NoiseLevelApplied = int(rec["noise_level"])
        self.TotalCoverage = int(rec["total_coverage"])
        self.EstimatedCoverageByDirection = [int(x) for x in rec["coverage_by_direction"]]
        self.SupportByDirection = [int(x) for x in rec["support_by_direction"]]
        self.AlleleSupport = int(rec["allele_support"])
        self.ReferenceSupport = int(rec["reference_support"])
        self.NumNoCalls = int(rec["num_no_calls"])
        self.FractionNoCalls = float(rec["fraction_no_calls"])
        self.SumOfBaseQuality = float(rec["sum_base_quality"])
        self.BiasScore = float(rec["bias_score"])
        self.GATKBiasScore = float(rec["gatk_bias_score"])
        f = int(rec["sb_flags"])
        self.BiasAcceptable, self.VarPresentOnBothStrands, self.CovPresentOnBothStrands = bool(f & 1), bool(f & 2), bool(f & 4)
        rl, al, ab = int(rec["ref_len"]), int(rec["alt_len"]), int(rec["allele_bytes"])
        raw = ab.to_bytes(4, "little") if rl + al <= 4 else bytes(arena[ab:ab + rl + al])
        self.ReferenceAllele = raw[:rl].decode()
        self.AlternateAllele = raw[rl:rl + al].decode()

    @property
    def Frequency(self):  # CalledAllele.cs:49-52 (float32 arithmetic)
        if self.TotalCoverage == 0:
            return np.float32(0)
        return min(np.float32(self.AlleleSupport) / np.float32(self.TotalCoverage), np.float32(1))

    def __repr__(self):
        return (f"<{self.Type.name} {self.ReferencePosition} {self.ReferenceAllele}>{self.AlternateAllele} Q{self.VariantQscore} "
                f"GQ{self.GenotypeQscore} {self.Genotype.name} DP{self.TotalCoverage} AD{self.AlleleSupport} {[x.name for x in self.Filters]}>")


class GpuStateManager:
    """IStateManager over libpisces_b200.so: buffers reads / pileups on the device; counts live in HBM until called."""

    def __init__(self, config=None, chr_name="chr1", chr_sequence=None, intervals=None):
        self._L = N.load()
        self.config = config if config is not None else make_config()
        h = C.c_void_p()
        rc = self._L.pb2_create(C.byref(self.config), C.byref(h))
        if rc != 0:
            raise PiscesB200Error(rc, self._L.pb2_last_error(None).decode())
        self._h = h
        self._keep = []
        if chr_sequence is not None:
            seq = chr_sequence.encode() if isinstance(chr_sequence, str) else bytes(chr_sequence)
            self._chk(self._L.pb2_set_reference(self._h, chr_name.encode(), seq, len(seq)))
        if intervals is not None:
            s = np.ascontiguousarray([a for a, _ in intervals], dtype=np.int32)
            e = np.ascontiguousarray([b for _, b in intervals], dtype=np.int32)
            self._chk(self._L.pb2_set_intervals(self._h, s.ctypes.data, e.ctypes.data, len(s)))

    # -- plumbing
    def _chk(self, rc):
        if rc != 0:
            raise PiscesB200Error(rc, self._L.pb2_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.pb2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def TotalNumCollapsed(self):
        n = C.c_int64()
        self._chk(self._L.pb2_totals(self._h, C.byref(n)))
        return n.value

    @property
    def ExpectStitchedReads(self):
        return bool(self.config.expect_stitched)

    # -- IStateManager
    def AddAlleleCounts(self, reads):
        """IStateManager.AddAlleleCounts(Read) (+ the SNV part of ICandidateVariantFinder.FindCandidates); accepts one Read or a list."""
        if isinstance(reads, Read):
            reads = [reads]
        if not reads:
            return
        n = len(reads)
        pos0 = np.array([r.position - 1 for r in reads], dtype=np.int32)
        flag = np.array([r.flag for r in reads], dtype=np.uint16)
        cig_off = np.zeros(n + 1, dtype=np.int64)
        seq_off = np.zeros(n + 1, dtype=np.int64)
        for i, r in enumerate(reads):
            cig_off[i + 1] = cig_off[i] + len(r.cigar)
            seq_off[i + 1] = seq_off[i] + len(r.sequence)
        cigar = np.array([c for r in reads for c in r.cigar], dtype=np.uint32)
        bases = np.frombuffer("".join(r.sequence for r in reads).encode(), dtype=np.uint8)
        quals = np.array([q for r in reads for q in r.qualities], dtype=np.uint8)
        has_dirs = any(r.base_directions is not None for r in reads)
        dirs = None
        if has_dirs:
            dirs = np.array([d for r in reads for d in (r.base_directions if r.base_directions is not None
                                                        else [1 if r.flag & 0x10 else 0] * len(r.sequence))], dtype=np.uint8)
        has_coll = any(r.collapsed is not None for r in reads)
        coll = np.array([r.collapsed or 0 for r in reads], dtype=np.uint8) if has_coll else None
        b = N.ReadBatch(n, pos0.ctypes.data, flag.ctypes.data, cig_off.ctypes.data, cigar.ctypes.data if len(cigar) else None,
                        seq_off.ctypes.data, bases.ctypes.data if len(bases) else None, quals.ctypes.data if len(quals) else None,
                        dirs.ctypes.data if dirs is not None else None, coll.ctypes.data if coll is not None else None)
        if len(cigar) == 0 or len(bases) == 0:
            raise PiscesB200Error(N_ERR_ARG, "reads without CIGAR or bases")
        self._chk(self._L.pb2_push_reads(self._h, C.byref(b)))

    def AddPileup(self, offsets, code, qual, anchor, first_position=1, positions=None, ref_bases=None, device=False):
        """Stage a locus-major pileup (pb2_push_pileup / pb2_push_pileup_device). Arrays are numpy (host) or torch CUDA tensors (device)."""
        def ptr(a):
            if a is None:
                return None
            return a.data_ptr() if device else np.ascontiguousarray(a).ctypes.data
        if not device:
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            code, qual, anchor = (np.ascontiguousarray(x, dtype=np.uint8) for x in (code, qual, anchor))
            if positions is not None:
                positions = np.ascontiguousarray(positions, dtype=np.int32)
            if ref_bases is not None:
                ref_bases = np.ascontiguousarray(ref_bases, dtype=np.uint8)
        n_loci = (offsets.numel() if device else len(offsets)) - 1
        p = N.PileupCsr(n_loci, int(first_position), ptr(positions), ptr(offsets), ptr(code), ptr(qual), ptr(anchor), ptr(ref_bases))
        self._keep = [offsets, code, qual, anchor, positions, ref_bases]
        self._chk((self._L.pb2_push_pileup_device if device else self._L.pb2_push_pileup)(self._h, C.byref(p)))

    @staticmethod
    def pack_pileup(code, qual, anchor, offsets=None, ref_bases=None):
        """Host planes (code, qual, anchor) -> PB2_LAYOUT_PACKED2: two bytes per entry plus the sparse list of candidate flags. With offsets and
        ref_bases given, flags on entries whose base is the reference base of their locus are left out: such a base raises no SNV candidate
        (CandidateVariantFinder.cs:112-141), so its open-end flags say nothing (staging clears them as well)."""
        L = N.load()   # pb2_pack_pileup: the packing is library code (a .NET host calls the same entry point)
        code, qual, anchor = (np.ascontiguousarray(x, dtype=np.uint8) for x in (code, qual, anchor))
        n = len(code)
        by_locus = offsets is not None and ref_bases is not None
        off = np.ascontiguousarray(offsets, dtype=np.int64) if by_locus else None
        rb = np.ascontiguousarray(ref_bases, dtype=np.uint8) if by_locus else None
        pcode, pqual = np.empty(n, dtype=np.uint8), np.empty(n, dtype=np.uint8)
        cap = max(1024, n // 64)
        while True:
            fi, fb = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.uint8)
            nf = L.pb2_pack_pileup(code.ctypes.data, qual.ctypes.data, anchor.ctypes.data, n, off.ctypes.data if by_locus else None, rb.ctypes.data if by_locus else None,
                                   len(off) - 1 if by_locus else 0, pcode.ctypes.data, pqual.ctypes.data, fi.ctypes.data, fb.ctypes.data, cap)
            if nf == N_ERR_UNSUPPORTED:
                raise ValueError("PB2_LAYOUT_PACKED2 cannot carry the collapsed-read type (anchor bits 4-7)")
            if nf < 0:
                raise PiscesB200Error(int(nf), "pb2_pack_pileup")
            if nf <= cap:
                break
            cap = int(nf)
        return pcode, pqual, fi[:nf].copy(), fb[:nf].copy()

    def AddPileupPacked(self, offsets, pcode, pqual, flag_index=None, flag_bits=None, first_position=1, positions=None, ref_bases=None):
        """pb2_push_pileup with PB2_LAYOUT_PACKED2 host buffers (see pack_pileup): 2 bytes per entry cross the PCIe link instead of 3."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        pcode, pqual = (np.ascontiguousarray(x, dtype=np.uint8) for x in (pcode, pqual))
        nf = 0 if flag_index is None else len(flag_index)
        fi = np.ascontiguousarray(flag_index, dtype=np.int64) if nf else None
        fb = np.ascontiguousarray(flag_bits, dtype=np.uint8) if nf else None
        if positions is not None:
            positions = np.ascontiguousarray(positions, dtype=np.int32)
        if ref_bases is not None:
            ref_bases = np.ascontiguousarray(ref_bases, dtype=np.uint8)
        ptr = lambda a: None if a is None else a.ctypes.data
        p = N.PileupCsr(len(offsets) - 1, int(first_position), ptr(positions), ptr(offsets), ptr(pcode), ptr(pqual), None, ptr(ref_bases), 1, 0, nf, ptr(fi), ptr(fb))
        self._keep = [offsets, pcode, pqual, fi, fb, positions, ref_bases]
        self._chk(self._L.pb2_push_pileup(self._h, C.byref(p)))

    def AddReadsSoA(self, d):
        """pb2_push_reads with a struct of arrays (dict of numpy arrays / pinned torch tensors: pos0 int32, flag uint16, cigar_off int64, cigar uint32,
        seq_off int64, bases, quals[, base_dirs, collapsed, amplicon int32]): IStateManager.AddAlleleCounts + FindCandidates for every read of the batch."""
        def ptr(a, dt):
            if a is None:
                return None
            if hasattr(a, "data_ptr"):       # torch tensor (pinned host memory)
                return a.data_ptr()
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data
        keep = []
        b = N.ReadBatch(int(len(d["pos0"])), ptr(d["pos0"], np.int32), ptr(d["flag"], np.uint16), ptr(d["cigar_off"], np.int64), ptr(d["cigar"], np.uint32),
                        ptr(d["seq_off"], np.int64), ptr(d["bases"], np.uint8), ptr(d["quals"], np.uint8), ptr(d.get("base_dirs"), np.uint8),
                        ptr(d.get("collapsed"), np.uint8), ptr(d.get("amplicon"), np.int32))
        self._chk(self._L.pb2_push_reads(self._h, C.byref(b)))

    @staticmethod
    def pack_reads(d, compact=False):
        """pb2_pack_reads: the struct of arrays of AddReadsSoA with bases + quals replaced by one packed byte per base (+ the exception list).
        compact: also the compact-offsets form (one byte of operation count per read instead of two 8-byte offsets; the offsets are built on the device)."""
        L = N.load()
        if compact:
            co, so = np.asarray(d["cigar_off"], dtype=np.int64), np.asarray(d["seq_off"], dtype=np.int64)
            ops = np.diff(co)
            if len(ops) and ops.max() > 255:
                raise ValueError("compact offsets need at most 255 CIGAR operations per read")
            d = dict(d, cigar=np.asarray(d["cigar"])[co[0]:co[-1]], bases=np.asarray(d["bases"])[so[0]:so[-1]], quals=np.asarray(d["quals"])[so[0]:so[-1]],
                     base_dirs=None if d.get("base_dirs") is None else np.asarray(d["base_dirs"])[so[0]:so[-1]])
        bases, quals = np.ascontiguousarray(d["bases"], dtype=np.uint8), np.ascontiguousarray(d["quals"], dtype=np.uint8)
        n = len(bases)
        seq = np.empty(n, dtype=np.uint8)
        cap = max(1024, n // 64)
        while True:
            ei, eb, eq = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.uint8), np.empty(cap, dtype=np.uint8)
            ne = L.pb2_pack_reads(bases.ctypes.data, quals.ctypes.data, n, seq.ctypes.data, ei.ctypes.data, eb.ctypes.data, eq.ctypes.data, cap)
            if ne < 0:
                raise PiscesB200Error(int(ne), "pb2_pack_reads")
            if ne <= cap:
                break
            cap = int(ne)
        out = {k: d[k] for k in ("pos0", "flag", "cigar_off", "cigar", "seq_off")}
        if compact:
            out.update(cigar_off=None, seq_off=None, cigar_ops=ops.astype(np.uint8), n_cigar_total=int(len(d["cigar"])), n_seq_total=int(n))
        out.update(seq=seq, exc_index=ei[:ne].copy(), exc_base=eb[:ne].copy(), exc_qual=eq[:ne].copy(), base_dirs=d.get("base_dirs"), collapsed=d.get("collapsed"), amplicon=d.get("amplicon"))
        return out

    def AddReadsPacked(self, d):
        """pb2_push_reads_packed with the dict pack_reads returns (numpy arrays or pinned torch tensors)."""
        keep = []

        def ptr(a, dt):
            if a is None:
                return None
            if hasattr(a, "data_ptr"):
                return a.data_ptr()
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data if len(a) else None
        b = N.PackedReadBatch(int(len(d["pos0"])), ptr(d["pos0"], np.int32), ptr(d["flag"], np.uint16), ptr(d.get("cigar_off"), np.int64), ptr(d["cigar"], np.uint32),
                              ptr(d.get("seq_off"), np.int64), ptr(d["seq"], np.uint8), int(len(d["exc_index"])), ptr(d["exc_index"], np.int64), ptr(d["exc_base"], np.uint8),
                              ptr(d["exc_qual"], np.uint8), ptr(d.get("base_dirs"), np.uint8), ptr(d.get("collapsed"), np.uint8), ptr(d.get("amplicon"), np.int32),
                              ptr(d.get("cigar_ops"), np.uint8), int(d.get("n_cigar_total") or 0), int(d.get("n_seq_total") or 0))
        self._chk(self._L.pb2_push_reads_packed(self._h, C.byref(b)))

    def AddReadBatch(self, batch):
        """pb2_push_reads / pb2_push_reads_packed with a ready batch struct (e.g. from BamReadStager): IStateManager.AddAlleleCounts + FindCandidates for
        each read."""
        push = self._L.pb2_push_reads_packed if isinstance(batch, N.PackedReadBatch) else self._L.pb2_push_reads
        self._chk(push(self._h, C.byref(batch)))

    def AddCandidates(self, candidates, arena=None):
        """IAlleleSource.AddCandidates (pb2_push_candidates). Either a list of dicts (type, pos, ref, alt, support[3], well_anchored[3], open_left,
        open_right, collapsed_mut[8]) or a numpy array of pb2_candidate rows plus the allele arena they point into."""
        if arena is None:
            arr = np.zeros(len(candidates), dtype=N.CANDIDATE_DTYPE)
            buf = bytearray()
            for i, c in enumerate(candidates):
                r = arr[i]
                r["position"], r["type"] = c["pos"], int(c["type"])
                r["open_flags"] = (1 if c.get("open_left") else 0) | (2 if c.get("open_right") else 0)
                r["ref_len"], r["alt_len"], r["allele_offset"] = len(c["ref"]), len(c["alt"]), len(buf)
                buf += c["ref"].encode() + c["alt"].encode()
                r["support"] = tuple(c.get("support", (0, 0, 0)))
                r["well_anchored"] = tuple(c.get("well_anchored", (0, 0, 0)))
                r["collapsed_mut"] = tuple(c.get("collapsed_mut", (0,) * 8))
            candidates, arena = arr, bytes(buf)
        candidates = np.ascontiguousarray(candidates)
        assert candidates.dtype.itemsize == 72
        ab = np.frombuffer(arena, dtype=np.uint8) if len(arena) else np.zeros(1, dtype=np.uint8)
        self._chk(self._L.pb2_push_candidates(self._h, candidates.ctypes.data, len(candidates), ab.ctypes.data, len(arena)))

    def SetForcedAlleles(self, alleles):
        """The forcedGtAlleles of Factory.CreateSomaticVariantCaller / SmallVariantCaller's constructor (SmallVariantCaller.cs:48-77) for this chromosome:
        an iterable of (position, ref, alt). pb2_set_forced_alleles."""
        alleles = list(alleles)
        arr = np.zeros(max(len(alleles), 1), dtype=N.CANDIDATE_DTYPE)
        buf = bytearray()
        for i, (pos, ref, alt) in enumerate(alleles):
            r = arr[i]
            r["position"], r["ref_len"], r["alt_len"], r["allele_offset"] = pos, len(ref), len(alt), len(buf)
            buf += ref.encode() + alt.encode()
        ab = np.frombuffer(bytes(buf), dtype=np.uint8) if len(buf) else np.zeros(1, dtype=np.uint8)
        self._chk(self._L.pb2_set_forced_alleles(self._h, arr.ctypes.data, len(alleles), ab.ctypes.data, len(buf)))

    def FormatVcf(self, records, ext=None, debug_mode=False, output_bias_files=False, report_rc_counts=False, report_ts_counts=False, crushed=False,
                  report_no_calls=False, pad_intervals=0):
        """pb2_vcf_format: the VCF record lines of raw records (Call(..., raw=True)) of the last flush, as a list of strings. crushed: one line per
        position (the germline writer, VcfFileWriter.GroupsAllelesThenWrite); pad_intervals: RegionMapper's empty reference calls (1 = before each
        written position, 2 = also after the last one)."""
        records = np.ascontiguousarray(records)
        opt = (C.c_int32 * 8)(int(debug_mode), int(output_bias_files), int(report_rc_counts), int(report_ts_counts), int(crushed), int(report_no_calls),
                              int(pad_intervals), 0)
        text, n = C.c_char_p(), C.c_int64()
        e = None if ext is None else np.ascontiguousarray(ext)
        self._chk(self._L.pb2_vcf_format(self._h, records.ctypes.data, None if e is None else e.ctypes.data, len(records), opt, C.byref(text), C.byref(n)))
        return C.string_at(text, n.value).decode().split("\n")[:-1]

    def AlleleArena(self):
        """Bytes that pb2_call_record.allele_bytes of the last Call points into for alleles longer than 4 bases."""
        p, n = C.c_void_p(), C.c_int64()
        self._chk(self._L.pb2_allele_arena(self._h, C.byref(p), C.byref(n)))
        return C.string_at(p.value, n.value) if n.value else b""

    def AlleleExt(self):
        """pb2_call_record_ext rows (ReadCollapsedCountsMut / ...Total, WellAnchoredSupportByDirection) of the records of the last Call, same order."""
        p, n = C.c_void_p(), C.c_int64()
        self._chk(self._L.pb2_flush_ext(self._h, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=N.RECORD_EXT_DTYPE)
        return np.frombuffer((C.c_char * (80 * n.value)).from_address(p.value), dtype=N.RECORD_EXT_DTYPE).copy()

    def GetAlleleCounts(self, position0, n):
        """RegionState._alleleCounts over [position0, position0+n): int32 [n][6][3][11]."""
        out = np.zeros((n, 6, 3, 11), dtype=np.int32)
        self._chk(self._L.pb2_get_counts(self._h, position0, n, out.ctypes.data))
        return out

    def GetAlleleCount(self, position, alleleType, directionType, minAnchor=0, maxAnchor=None, fromEnd=False, symmetric=False):
        """IAlleleSource.GetAlleleCount (anchor selection per AlleleCountHelper.cs:21-85, evaluated on the 11 device-computed bins)."""
        bins = self.GetAlleleCounts(position, 1)[0, int(alleleType), int(directionType)]
        K, NA = 5, 11
        true_min = min(K, minAnchor)
        init_max = K
        if maxAnchor is not None:
            init_max = K - 1 if maxAnchor >= K else maxAnchor
        tot = 0
        if fromEnd:
            tot += sum(int(bins[NA - i - 1]) for i in range(true_min, init_max + 1))
            if maxAnchor is None:
                tot += sum(int(bins[i]) for i in range(true_min if symmetric else 0, init_max))
        else:
            tot += sum(int(bins[i]) for i in range(true_min, init_max + 1))
            if maxAnchor is None:
                tot += sum(int(bins[i]) for i in range(init_max + 1, NA - true_min if symmetric else NA))
        return tot

    def DoneProcessing(self):
        self._chk(self._L.pb2_reset(self._h))

    # -- used by GpuAlleleCaller
    def _flush(self, up_to, copy=True):
        out = C.c_void_p()
        n = C.c_int64()
        self._chk(self._L.pb2_flush(self._h, -1 if up_to is None else int(up_to), C.byref(out), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=N.RECORD_DTYPE)
        buf = (C.c_char * (96 * n.value)).from_address(out.value)
        view = np.frombuffer(buf, dtype=N.RECORD_DTYPE)
        return view.copy() if copy else view   # the view is the library's own host buffer: valid until the next flush / reset / close of this handle

    def StageReads(self):
        """pb2_stage_reads: everything pushed through AddAlleleCounts / AddReadBatch becomes one device-resident segment for call_resident."""
        self._chk(self._L.pb2_stage_reads(self._h))

    def stage_stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_double()
        self._chk(self._L.pb2_stage_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(staged_bytes=a.value, rows=b.value, stage_ms=c.value)

    def flush_resident(self, copy=True):
        """pb2_flush_resident: the whole job from the device-resident reads (find candidates, stage, call), nothing consumed. Returns raw records
        (copy=False: a view of the library's host buffer, valid until the next flush)."""
        out = C.c_void_p()
        n = C.c_int64()
        self._chk(self._L.pb2_flush_resident(self._h, C.byref(out), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=N.RECORD_DTYPE)
        view = np.frombuffer((C.c_char * (96 * n.value)).from_address(out.value), dtype=N.RECORD_DTYPE)
        return view.copy() if copy else view

    def SetOwnedRange(self, own_lo, own_hi):
        """pb2_set_owned_range: this handle is one interval shard and emits only the positions it owns."""
        self._chk(self._L.pb2_set_owned_range(self._h, int(own_lo), int(own_hi)))

    def set_resident_sink(self, device_ptr, slot_records, n_slots):
        self._chk(self._L.pb2_set_resident_sink(self._h, device_ptr, int(slot_records), int(n_slots)))

    def call_resident_async(self):
        self._chk(self._L.pb2_call_resident_async(self._h))

    def resident_sync(self):
        n = C.c_int64()
        self._chk(self._L.pb2_resident_sync(self._h, C.byref(n)))
        return n.value

    def sink_sort(self):
        self._chk(self._L.pb2_sink_sort(self._h))

    def call_resident(self):
        n = C.c_int64()
        self._chk(self._L.pb2_call_resident(self._h, C.byref(n)))
        return n.value

    def stats(self):
        a, b, c = C.c_int64(), C.c_double(), C.c_int64()
        self._chk(self._L.pb2_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(hot_launches=a.value, hot_ms=b.value, total_launches=c.value)


N_ERR_ARG = -1
N_ERR_UNSUPPORTED = -5


class BamReadStager:
    """pb2_bam_*: BAM file -> pb2_read_batch batches (BGZF inflate, record decode, AlignmentSource.ShouldSkipRead, XD / XV / XW / XR), all in the library.
    Iterating yields (ref_id, ReadBatch struct, n_skipped); a batch is only valid until the next one is fetched."""

    def __init__(self, path, min_map_quality=1, remove_duplicates=True, only_proper_pairs=False, max_reads=65536, packed=False):
        self._L = N.load()
        self.packed = packed   # hand out pb2_packed_read_batch (pb2_bam_next_batch_packed: one byte per base, compact offsets) instead of pb2_read_batch
        self._r = C.c_void_p()
        if self._L.pb2_bam_open(os.fsencode(path), C.byref(self._r)) != 0:
            raise PiscesB200Error(N_ERR_ARG, f"pb2_bam_open({path}) failed")
        self._flt = (C.c_int32 * 3)(int(min_map_quality), int(remove_duplicates), int(only_proper_pairs))
        self.max_reads = max_reads
        n, names, lens, st, co = C.c_int32(), C.POINTER(C.c_char_p)(), C.POINTER(C.c_int32)(), C.c_int32(), C.c_int32()
        self._L.pb2_bam_header(self._r, C.byref(n), C.byref(names), C.byref(lens), C.byref(st), C.byref(co))
        self.references = [(names[i].decode(), int(lens[i])) for i in range(n.value)]
        self.is_stitched, self.is_collapsed = bool(st.value), bool(co.value)

    def __iter__(self):
        while True:
            b, ref_id, skipped = (N.PackedReadBatch() if self.packed else N.ReadBatch()), C.c_int32(), C.c_int64()
            nxt = self._L.pb2_bam_next_batch_packed if self.packed else self._L.pb2_bam_next_batch
            if nxt(self._r, self._flt, self.max_reads, C.byref(b), C.byref(ref_id), C.byref(skipped)) != 0:
                raise PiscesB200Error(N_ERR_ARG, self._L.pb2_bam_last_error(self._r).decode())
            if b.n_reads == 0:
                return
            yield ref_id.value, b, skipped.value

    def batch_amplicons(self):
        """Read.GetAmpliconNameIfExists of the reads of the batch fetched last: ids into amplicon_names() (-1 = no XN tag). pb2_bam_batch_amplicons."""
        p, n = C.POINTER(C.c_int32)(), C.c_int32()
        if self._L.pb2_bam_batch_amplicons(self._r, C.byref(p), C.byref(n)) != 0:
            raise PiscesB200Error(N_ERR_ARG, "pb2_bam_batch_amplicons failed")
        return [int(p[i]) for i in range(n.value)]

    def amplicon_names(self):
        """The file's amplicon names so far, in first-seen order over the kept reads. pb2_bam_amplicon_names."""
        n, names = C.c_int32(), C.POINTER(C.c_char_p)()
        if self._L.pb2_bam_amplicon_names(self._r, C.byref(n), C.byref(names)) != 0:
            raise PiscesB200Error(N_ERR_ARG, "pb2_bam_amplicon_names failed")
        return [names[i].decode() for i in range(n.value)]

    def close(self):
        if self._r:
            self._L.pb2_bam_close(self._r)
            self._r = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuAlleleCaller:
    """IAlleleCaller: Call(batch, source) -> {position: [CalledAllele, ...]} (SortedList<int, List<CalledAllele>>)."""

    def __init__(self):
        self.TotalNumCalled = 0
        self.TotalNumCollapsed = 0

    def Call(self, source, upToPosition=None, raw=False, copy=True):
        """raw: the pb2_call_record array instead of CalledAllele objects; with copy=False a view of the library's host buffer (what a P/Invoke host
        reads in place), valid until the handle's next flush."""
        recs = source._flush(upToPosition, copy=copy or not raw)
        self.TotalNumCalled += len(recs)
        if raw:
            return recs
        out = {}
        arena = source.AlleleArena()
        for r in recs:
            out.setdefault(int(r["position"]), []).append(CalledAllele(r, arena))
        return out
