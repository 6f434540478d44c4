"""Parity of the explicit-candidate path (insertions, deletions, MNVs: pb2_push_candidates / the device candidate finder, spanning coverage,
collapser, MNV reallocator) against the CPU oracle and the reference's own CoverageCalculator vectors. Needs a B200: -m gpu."""
import math

import numpy as np
import pytest

from oracle import binding as ob
from tests import util_reads as U
from tests.util_counts import COVERAGE_VECTORS, pileup_from_counts

pytestmark = pytest.mark.gpu


def _pb():
    import pisces_b200 as pb
    return pb


def compare_records(orecs, precs, arena, check_qsum=False):
    assert len(orecs) == len(precs), (len(orecs), len(precs), [(o.pos, o.ref, o.alt) for o in orecs][:20],
                                      [(int(p["position"]), int(p["type"])) for p in precs][:20])
    for o, p in zip(orecs, precs):
        ctx = f"pos {o.pos} {o.ref}>{o.alt}"
        assert o.pos == int(p["position"]) and o.type == int(p["type"]), ctx
        rl, al, ab = int(p["ref_len"]), int(p["alt_len"]), int(p["allele_bytes"])
        raw = ab.to_bytes(4, "little") if rl + al <= 4 else bytes(arena[ab:ab + rl + al])
        assert raw[:rl].decode() == o.ref and raw[rl:rl + al].decode() == o.alt, ctx
        assert o.total_coverage == int(p["total_coverage"]), ctx
        assert list(o.cov) == list(p["coverage_by_direction"]), ctx
        assert list(o.support) == list(p["support_by_direction"]), ctx
        assert o.allele_support == int(p["allele_support"]) and o.ref_support == int(p["reference_support"]), ctx
        assert o.num_no_calls == int(p["num_no_calls"]), ctx
        assert o.vq == int(p["variant_qscore"]) and o.gq == int(p["genotype_qscore"]), ctx
        assert o.genotype == int(p["genotype"]), ctx
        assert o.filter_mask == int(p["filters"]), (ctx, o.filter_mask, int(p["filters"]))
        assert o.noise_level == int(p["noise_level"]), ctx
        assert o.fraction_no_calls == float(p["fraction_no_calls"]), ctx
        assert (bool(o.bias_acceptable), bool(o.var_both_strands), bool(o.cov_both_strands), bool(o.forced)) == \
            (bool(p["sb_flags"] & 1), bool(p["sb_flags"] & 2), bool(p["sb_flags"] & 4), bool(p["sb_flags"] & 8)), ctx
        for a, b in ((o.bias_score, float(p["bias_score"])), (o.gatk_bias_score, float(p["gatk_bias_score"]))):
            if math.isinf(a) or math.isnan(a):
                assert (math.isinf(b) and (a > 0) == (b > 0)) or (math.isnan(a) and math.isnan(b)), ctx
            else:
                assert b == pytest.approx(a, rel=1e-6, abs=1e-9), ctx
        if check_qsum:
            assert float(p["sum_base_quality"]) == pytest.approx(o.sum_base_quality, rel=1e-9), ctx


@pytest.mark.parametrize("name", sorted(COVERAGE_VECTORS))
def test_reference_coverage_vectors_through_kernels(name):
    """CoverageCalculatorTests.cs (ComputeCoverage_Insertions / _Spanning_HappyPath): the staged counts become pileup entries, the allele an
    explicit candidate; the record's coverage must be the reference's expected values."""
    pb = _pb()
    v = COVERAGE_VECTORS[name]
    off, code, qual, anch = pileup_from_counts(v["counts"], v["n_loci"])
    sm = pb.GpuStateManager(pb.make_config(min_variant_qscore=0, min_frequency=1e-6, min_coverage=0, collapse=0, output_gvcf=0, expect_stitched=v.get("stitched", 0)),
                            "chr1", "A" * 64)
    sm.AddPileup(off, code, qual, anch, first_position=1)
    sm.AddCandidates([dict(type=v["type"], pos=1, ref=v["ref"], alt=v["alt"], support=v["support"], well_anchored=v["well_anchored"])])
    calls = pb.GpuAlleleCaller().Call(sm)
    sm.close()
    a = [c for c in calls[1] if int(c.Type) == v["type"]][0]
    assert a.EstimatedCoverageByDirection == list(v["expected_cov"]) and a.TotalCoverage == sum(v["expected_cov"])
    assert a.ReferenceSupport == max(0, a.TotalCoverage - a.AlleleSupport)
    assert (a.ReferenceAllele, a.AlternateAllele) == (v["ref"], v["alt"])


def _resident_records(sm):
    """Device-resident results of pb2_call_resident (dense reference stream + compacted variant stream), copied back with torch."""
    import ctypes as C
    import torch
    from pisces_b200 import _native as N

    class DevBuf:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    rr, rv, vr = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nl, nv = C.c_int64(), C.c_int64()
    sm._chk(sm._L.pb2_resident_results(sm._h, C.byref(rr), C.byref(rv), C.byref(nl), C.byref(vr), C.byref(nv)))
    out = []
    if nv.value:
        out.append(np.frombuffer(torch.as_tensor(DevBuf(vr.value, nv.value * 96), device="cuda").cpu().numpy().tobytes(), dtype=N.RECORD_DTYPE))
    if rr.value:
        refs = np.frombuffer(torch.as_tensor(DevBuf(rr.value, nl.value * 96), device="cuda").cpu().numpy().tobytes(), dtype=N.RECORD_DTYPE)
        valid = torch.as_tensor(DevBuf(rv.value, nl.value), device="cuda").cpu().numpy().astype(bool)
        out.append(refs[valid])
    return np.concatenate(out) if out else np.zeros(0, dtype=N.RECORD_DTYPE)


def _pileup_indels_both(d, gvcf, **kw):
    pb = _pb()
    ref = bytes(d["ref_bases"].numpy()).decode()
    off, code, qual, anch = (d[k].numpy() for k in ("offsets", "code", "qual", "anchor"))
    arena = d["arena"]
    oc = ob.Caller(ob.default_config(output_gvcf=gvcf, **kw), "chr1", ref)
    for c in d["candidates"]:
        o = int(c["allele_offset"])
        r, a = arena[o:o + int(c["ref_len"])].decode(), arena[o + int(c["ref_len"]):o + int(c["ref_len"]) + int(c["alt_len"])].decode()
        oc.add_candidate(int(c["type"]), int(c["position"]), r, a, [int(x) for x in c["support"]], [int(x) for x in c["well_anchored"]],
                         bool(c["open_flags"] & 1), bool(c["open_flags"] & 2))
    oc.add_pileup(off, code, qual, anch, 1)
    oc.finish()
    pkw = dict(kw)
    if pkw.get("noise_model") == 1:
        pkw["want_sum_base_quality"] = 1
    sm = pb.GpuStateManager(pb.make_config(output_gvcf=gvcf, **pkw), "chr1", ref)
    sm.AddPileup(off, code, qual, anch, first_position=1)
    sm.AddCandidates(d["candidates"], arena)
    # the resident (bench) path must produce the same records, on the first (plan-building) call and on a replayed one (explicit pass on the side stream)
    n_res = sm.call_resident()
    assert sm.call_resident() == n_res
    res = _resident_records(sm)
    precs = pb.GpuAlleleCaller().Call(sm, raw=True)
    parena = sm.AlleleArena()
    sm.close()
    key = lambda r: (int(r["position"]), int(r["type"]), int(r["ref_len"]), int(r["alt_len"]), int(r["allele_bytes"]) if int(r["ref_len"]) + int(r["alt_len"]) <= 4 else -1)
    assert sorted(bytes(r.tobytes()) for r in res if int(r["ref_len"]) + int(r["alt_len"]) <= 4) == \
        sorted(bytes(r.tobytes()) for r in precs if int(r["ref_len"]) + int(r["alt_len"]) <= 4)
    assert sorted(key(r) for r in res) == sorted(key(r) for r in precs)
    return oc.records(), precs, parena, n_res


@pytest.mark.parametrize("gvcf", [0, 1])
@pytest.mark.parametrize("depth,n_loci", [(60, 4000), (500, 3000)])
def test_locus_major_indels_match_oracle(depth, n_loci, gvcf):
    """BASELINE configs[1] shape (SNV + indel) at test size: synthetic insertions / deletions as explicit candidates next to the SNV / reference
    stream, against the oracle fed the same entries and candidates."""
    from pisces_b200 import synth
    d = synth.make_pileup(n_loci, depth, seed=depth + gvcf, snv_rate=0.03, indel_rate=0.02)
    orecs, precs, arena, n_res = _pileup_indels_both(d, gvcf)
    n_indel = sum(1 for o in orecs if o.type in (ob.INSERTION, ob.DELETION))
    assert n_indel > 20
    compare_records(orecs, precs, arena)
    assert n_res == sum(1 for o in orecs if o.type != ob.REFERENCE)


def test_locus_major_indels_window_noise_and_long_alleles():
    from pisces_b200 import synth
    d = synth.make_pileup(2000, 200, seed=9, snv_rate=0.02, indel_rate=0.03)
    orecs, precs, arena, _ = _pileup_indels_both(d, 1, noise_model=1)
    assert any(o.ref_len + o.alt_len > 4 for o in orecs)
    compare_records(orecs, precs, arena, check_qsum=True)


def test_indel_repeat_and_rmxn_filters():
    """A 1-base deletion inside a homopolymer run: IndelRepeatLength (AlleleProcessor.cs:80-213) and RMxN (RMxNCalculator.cs) fire on both sides."""
    pb = _pb()
    ref = "ACGTACGTTG" + "A" * 12 + "CGTACGTACGTTGCATGCAA"
    n = len(ref)
    counts = np.zeros((n, 6, 3, 11), dtype=np.int32)
    for i, b in enumerate(ref):
        a = {"A": 0, "G": 1, "C": 2, "T": 3}[b]
        counts[i, a, 0, 5] = 150
        counts[i, a, 1, 5] = 150
    counts[10, 0, 0, 5] -= 20
    counts[10, 5, 0, 5] += 20      # deletion of the first A of the run: position 10 (the G) is the anchor base
    off, code, qual, anch = pileup_from_counts(counts, n)
    cand = dict(type=2, pos=10, ref="GA", alt="G", support=(20, 0, 0), well_anchored=(20, 0, 0))
    kw = dict(indel_repeat_filter=8, output_gvcf=0, collapse=0)
    oc = ob.Caller(ob.default_config(**kw), "chr1", ref)
    oc.add_candidate(cand["type"], cand["pos"], cand["ref"], cand["alt"], cand["support"], cand["well_anchored"])
    oc.add_pileup(off, code, qual, anch, 1)
    oc.finish()
    sm = pb.GpuStateManager(pb.make_config(**kw), "chr1", ref)
    sm.AddPileup(off, code, qual, anch, first_position=1)
    sm.AddCandidates([cand])
    precs = pb.GpuAlleleCaller().Call(sm, raw=True)
    arena = sm.AlleleArena()
    sm.close()
    orecs = oc.records()
    assert len(orecs) == 1 and orecs[0].filter_mask & (1 << 7) and orecs[0].filter_mask & (1 << 9)
    compare_records(orecs, precs, arena)


def _reads_both(reads, ref, o_kw, p_kw, intervals=None, flush_every=None, forced=()):
    pb = _pb()
    oc = ob.Caller(ob.default_config(**o_kw), "chr1", ref, intervals=intervals)
    for f in forced:
        oc.add_forced(*f)
    for rd in reads:
        oc.add_read(U.to_oracle(rd))
    oc.finish()
    sm = pb.GpuStateManager(pb.make_config(**p_kw), "chr1", ref, intervals=intervals)
    if forced:
        sm.SetForcedAlleles(forced)
    caller = pb.GpuAlleleCaller()
    precs, arenas = [], []
    if flush_every:
        for i in range(0, len(reads), flush_every):
            chunk = reads[i:i + flush_every]
            sm.AddAlleleCounts([U.to_product(rd) for rd in chunk])
            r = caller.Call(sm, upToPosition=chunk[-1]["pos"] - 1, raw=True)
            precs.append((r, sm.AlleleArena()))
    else:
        sm.AddAlleleCounts([U.to_product(rd) for rd in reads])
    r = caller.Call(sm, raw=True)
    precs.append((r, sm.AlleleArena()))
    total_collapsed = sm.TotalNumCollapsed
    sm.close()
    return oc, precs, total_collapsed


def _compare_chunks(orecs, chunks):
    i = 0
    for recs, arena in chunks:
        compare_records(orecs[i:i + len(recs)], recs, arena)
        i += len(recs)
    assert i == len(orecs)


@pytest.mark.parametrize("collapse", [0, 1])
@pytest.mark.parametrize("seed", [41, 42])
def test_reads_with_indels_match_oracle(seed, collapse):
    """Reads with insertions, deletions and soft clips: candidates are found on the device (CandidateVariantFinder), merged per position,
    collapsed (open-ended indels at read ends) and scored with spanning coverage."""
    rng = np.random.default_rng(seed)
    ref = U.random_reference(rng, 2600)
    hot = {int(p): ("ACGT"[int(rng.integers(0, 4))], float(rng.uniform(0.05, 0.5))) for p in rng.integers(60, 2400, 30)}
    reads = U.make_reads(rng, ref, 6000, read_len=60, hotspots=hot, del_rate=0.08, ins_rate=0.08, clip_rate=0.1, indel_sites=40)
    kw = dict(output_gvcf=1, collapse=collapse)
    oc, chunks, ncoll = _reads_both(reads, ref, dict(kw, min_vq=10), dict(kw, min_variant_qscore=10))
    orecs = oc.records()
    assert sum(1 for o in orecs if o.type in (ob.INSERTION, ob.DELETION)) > 10
    _compare_chunks(orecs, chunks)
    # TotalNumCollapsed of the oracle also counts open-ended SNV merges, which the count-based SNV path resolves without materialising candidates
    assert (ncoll > 0) == bool(collapse) and ncoll <= oc.L.po_caller_total_collapsed(oc.h)


@pytest.mark.parametrize("seed", [51, 52])
def test_reads_with_mnvs_match_oracle(seed):
    """CallMNVs: the SNV/MNV state machine of the finder, collapsing, failed-MNV reallocation (incl. spill into the next block) and gapped-MNV
    reference take-away, streamed block by block like SmallVariantCaller.Execute."""
    rng = np.random.default_rng(seed)
    ref = U.random_reference(rng, 3300)
    hot = {}
    for p in rng.integers(60, 3100, 40):
        p = int(p)
        f = float(rng.uniform(0.03, 0.5))
        for k in range(int(rng.integers(1, 4))):
            hot[p + k * int(rng.integers(1, 3))] = ("ACGT"[int(rng.integers(0, 4))], f)
    for p in (998, 999, 1000, 1001, 1999, 2001):     # MNVs across the 1000-bp block boundaries
        hot[p] = ("ACGT"[int(rng.integers(0, 4))], 0.3)
    reads = U.make_reads(rng, ref, 9000, read_len=60, hotspots=hot, del_rate=0.03, ins_rate=0.03, clip_rate=0.1, linked_hotspots=True)
    kw = dict(output_gvcf=1, collapse=1, call_mnvs=1, max_size_mnv=3, max_gap_mnv=1)
    oc, chunks, ncoll = _reads_both(reads, ref, dict(kw, min_vq=10), dict(kw, min_variant_qscore=10), flush_every=700)
    orecs = oc.records()
    assert sum(1 for o in orecs if o.type == ob.MNV) > 5
    _compare_chunks(orecs, chunks)
    assert ncoll == oc.L.po_caller_total_collapsed(oc.h)


@pytest.mark.parametrize("gvcf", [0, 1])
@pytest.mark.parametrize("mnvs", [0, 1])
def test_forced_alleles_match_oracle(mnvs, gvcf):
    """Forced-genotyping alleles (SmallVariantCaller.cs:48-77,118-155; AlleleCaller.cs:98-118,143-150): alleles the reads support and alleles they do
    not, SNVs / indels (/ MNVs with CallMNVs), inside and outside the intervals, streamed block by block. Reported whether callable or not, with the
    reference allele of their position kept beside the ones that are only reported because they are forced."""
    rng = np.random.default_rng(70 + 2 * mnvs + gvcf)
    ref = U.random_reference(rng, 3300)
    hot = {int(p): ("ACGT"[int(rng.integers(0, 4))], float(rng.uniform(0.02, 0.5))) for p in rng.integers(60, 3100, 40)}
    reads = U.make_reads(rng, ref, 7000, read_len=60, hotspots=hot, del_rate=0.04, ins_rate=0.04, clip_rate=0.1, indel_sites=30, linked_hotspots=bool(mnvs))
    forced = []
    for p, (b, _) in list(hot.items())[:25]:        # alleles with support (some callable, some not)
        if ref[p - 1] != b:
            forced.append((p, ref[p - 1], b))
    for p in rng.integers(60, 3100, 25):             # alleles without any support, a few at uncovered or boundary positions
        p = int(p)
        alt = "ACGT"[("ACGT".index(ref[p - 1]) + 1 + int(rng.integers(0, 3))) % 4]
        forced.append((p, ref[p - 1], alt))
    for p in (1000, 1001, 2000, 3290):
        forced.append((p, ref[p - 1], "ACGT"[("ACGT".index(ref[p - 1]) + 1) % 4]))
    for p in rng.integers(100, 3000, 8):
        p = int(p)
        forced.append((p, ref[p - 1:p + 2], ref[p - 1]))              # a 2-base deletion
        forced.append((p + 7, ref[p + 6], ref[p + 6] + "GT"))        # a 2-base insertion
    if mnvs:
        for p in rng.integers(100, 3000, 8):
            p = int(p)
            forced.append((p, ref[p - 1:p + 1], "".join("ACGT"[("ACGT".index(c) + 2) % 4] for c in ref[p - 1:p + 1])))
    forced = sorted(set(forced))
    if gvcf == 0 and mnvs == 1:     # the forced set's own (file) order decides which reference candidates GetClipped reaches: keep it unsorted once
        forced = [forced[int(i)] for i in rng.permutation(len(forced))]
    intervals = [(50, 1500), (1700, 3300)]
    kw = dict(output_gvcf=gvcf, collapse=1, call_mnvs=mnvs, max_size_mnv=3, max_gap_mnv=1)
    oc, chunks, _ = _reads_both(reads, ref, dict(kw, min_vq=20), dict(kw, min_variant_qscore=20), intervals=intervals, flush_every=900, forced=forced)
    orecs = oc.records()
    assert sum(1 for o in orecs if o.forced) > 20 and sum(1 for o in orecs if o.type == ob.REFERENCE) >= 1
    _compare_chunks(orecs, chunks)


@pytest.mark.parametrize("gvcf", [0, 1])
@pytest.mark.parametrize("seed", [61, 62, 63])
def test_streamed_reads_collapsable_snvs_from_later_blocks(seed, gvcf):
    """CallMNVs off, Collapse on, reads streamed block by block: when an indel reaches past the last block of a batch, the finished SNV candidates of
    the next block are pulled into it (RegionStateManager.AddCollapsableFromOtherBlocks :441-457) and RegionState.ExtractCollapsable (:470-490) removes them
    with List.Remove — the first candidate that Equals, open ends ignored — so an open-ended twin can be lost and the anchored one counted twice. The
    CUDA path turns the count-based SNVs of those positions into explicit candidates and replays exactly that."""
    rng = np.random.default_rng(seed)
    ref = U.random_reference(rng, 3300)
    hot = {int(p): ("ACGT"[int(rng.integers(0, 4))], float(rng.uniform(0.05, 0.5))) for p in rng.integers(60, 3100, 80)}
    reads = U.make_reads(rng, ref, 8000, read_len=60, hotspots=hot, del_rate=0.05, ins_rate=0.03, clip_rate=0.1, indel_sites=40, lowq_rate=0.12)
    for b in (1000, 2000, 3000):      # deletions that reach past the end of a 1000-bp block: MaxAlleleEndpoint > the block's end position
        for k in range(12):
            pos = b - 40 + k
            body = ref[pos - 1:b - 3] + ref[b + 3:b + 3 + 60 - (b - 2 - pos)]
            reads.append(dict(pos=pos, seq=body, cigar=f"{b - 2 - pos}M6D{60 - (b - 2 - pos)}M", quals=[35] * 60, flag=(16 if k % 2 else 0) | 67, dirs=None, coll=None,
                              xd=None, xr=None, xv=None, xw=None))
    reads.sort(key=lambda r: r["pos"])
    kw = dict(output_gvcf=gvcf, collapse=1)
    oc, chunks, _ = _reads_both(reads, ref, dict(kw, min_vq=20), dict(kw, min_variant_qscore=20), flush_every=900)
    oc0 = ob.Caller(ob.default_config(min_vq=20, output_gvcf=gvcf, collapse=0), "chr1", ref)
    for rd in reads:
        oc0.add_read(U.to_oracle(rd))
    oc0.finish()
    plain = {(o.pos, o.alt): list(o.support) for o in oc0.records() if o.type == ob.SNV}
    orecs = oc.records()
    # the scenario really occurs: some SNV's support differs from the plain count of its base
    assert any(o.type == ob.SNV and plain.get((o.pos, o.alt)) not in (None, list(o.support)) for o in orecs)
    _compare_chunks(orecs, chunks)


@pytest.mark.parametrize("min_vq,forced,golden", [(1, False, "phix_s3_noisy.records.vcf"), (1, True, "phix_s3_forced1.records.vcf"),
                                                  (20, True, "phix_s3_forced2.records.vcf")])
def test_phix_full_text_golden_through_cuda_path(min_vq, forced, golden):
    """The reference's own full-text goldens (PhiX_S3.bam -> PhiX_S3.noisy.vcf, .Forced1.vcf, .Forced2.vcf; ForcedGTFxnlTest.cs:11-112) reproduced
    line by line from the records the CUDA path emits: MNVs up to 10 with gaps of 5, collapsing, MNV reallocation, gapped-MNV reference take-away,
    q-scores, strand bias, the forced alleles of PhiX_S3.forcedGTInput.vcf (pb2_set_forced_alleles), and the VCF text of the library's writer (pb2_vcf_format)."""
    import gzip
    import json
    import os
    pb = _pb()
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    d = json.load(gzip.open(os.path.join(G, "phix_s3_reads.json.gz"), "rt"))
    genome = open(os.path.join(G, "phix_genome.txt")).read().strip()
    pkw = dict(min_coverage=2, min_base_call_quality=10, min_variant_qscore=min_vq, min_frequency=0.00001, forced_noise_level=40, call_mnvs=1, max_size_mnv=10,
               max_gap_mnv=5, no_call_filter=1.0)
    # AlignmentSource.ShouldSkipRead (AlignmentsSource.cs:84-92): mapped, primary, not duplicate, mapq >= 1, has a CIGAR — host-side filter
    reads = [r for r in d["reads"] if not (r["flag"] & 0x4) and not (r["flag"] & 0x100) and not (r["flag"] & 0x400) and r["mapq"] >= 1 and r["cigar"]]
    sm = pb.GpuStateManager(pb.make_config(**pkw), "phix", genome)
    if forced:   # Factory.cs:80-93: forced alleles with a non-ACGT alternate (".") are dropped by the host
        fa = json.load(open(os.path.join(G, "phix_forced_alleles.json")))
        sm.SetForcedAlleles([(p, r, a) for p, r, a in fa if all(ch in "ACGT" for ch in a)])
    caller = pb.GpuAlleleCaller()
    got = []      # the VCF text comes from the library's own writer (pb2_vcf_format), formatted flush by flush
    for i in range(0, len(reads), 50):      # streamed like SmallVariantCaller.Execute: push a few reads, call up to the last read's position - 1
        chunk = reads[i:i + 50]
        sm.AddAlleleCounts([pb.Read(r["pos0"] + 1, r["seq"], r["cigar"], r["qual"], flag=r["flag"]) for r in chunk])
        got += sm.FormatVcf(caller.Call(sm, upToPosition=chunk[-1]["pos0"], raw=True))
    got += sm.FormatVcf(caller.Call(sm, raw=True))
    sm.close()
    exp = [l.rstrip("\n") for l in open(os.path.join(G, golden))]
    assert len(got) == len(exp)
    for a, b in zip(got, exp):
        assert a == b


def test_collapsed_stitched_full_text_golden_through_cuda_path():
    """src/test/Pisces.Tests/FunctionalTests/SomaticVariantCallerFunctionalTests.cs:683-758: collapsed.test.stitched.bam against the mock chr1 ->
    test_truth.stitched.genome.vcf (171 records), line by line from the CUDA path's records: stitched directions (XD), collapsed-read categories
    (XV/XW/XR), MNVs up to 100 with gaps of 10, and the 12-field US tag from pb2_call_record_ext. The reference test builds its options by hand
    (no Validate()), hence skip_validation."""
    import json
    import os
    pb = _pb()
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    reads = json.load(open(os.path.join(G, "collapsed_stitched_reads.json")))
    seq = "N" * (9770498 - 1) + ("GAAGTAACAACGCAGGATGCCCCCTGGGGTGGACTGCCCCATGGAATTCTGGACCAAGGAGGAGAATCAGAGCGTTGTGGTTGACTTCCTGCTGCCCACAGGGGTCTACCTGAACTTCCCTGTGTCCCGCAATGCCAACCTC"
                                 "AGCACCATCAAGCAGGTATGGCCTCCATC")
    # AmpliconBiasFilterThreshold = 0.01F as the reference's test sets it (:719): tracking is on, the file has no XN tag, nothing is filtered
    pkw = dict(call_mnvs=1, max_size_mnv=100, max_gap_mnv=10, expect_stitched=1, expect_collapsed=1, skip_validation=1, amplicon_bias_filter=0.01)

    def base_dirs(r):   # Read.SequencedBaseDirectionMap: the XD runs projected onto the read bases (Read.cs:390-421,664-682)
        exp, num = [], ""
        for ch in r["xd"]:
            if ch.isdigit():
                num += ch
            else:
                exp += [{"F": 0, "R": 1, "S": 2}[ch]] * int(num)
                num = ""
        dirs, ci = [], 0
        for c in r["cigar"]:
            op, ln = c & 15, c >> 4
            for _ in range(ln):
                if op in (0, 1, 4, 7, 8):
                    dirs.append(exp[ci])
                ci += 1
        return dirs

    def collapsed_byte(r):   # pb2_read_batch.collapsed: IsCollapsedRead, IsDuplex, ReadPairDirection (Read.cs:17-71,311-349)
        has = r["xv"] is not None or r["xw"] is not None
        duplex = bool(r["xv"]) and bool(r["xw"])
        return (1 if has else 0) | (2 if duplex else 0) | ({"FR": 1, "RF": 2}.get(r["xr"], 0) << 2)

    sm = pb.GpuStateManager(pb.make_config(**pkw), "chr1", seq)
    sm.AddAlleleCounts([pb.Read(r["pos0"] + 1, r["seq"], r["cigar"], r["qual"], flag=r["flag"], base_directions=base_dirs(r), collapsed=collapsed_byte(r)) for r in reads])
    recs = pb.GpuAlleleCaller().Call(sm, raw=True)
    ext = sm.AlleleExt()
    assert len(ext) == len(recs)
    got = sm.FormatVcf(recs, ext, debug_mode=True, output_bias_files=True, report_rc_counts=True, report_ts_counts=True)
    sm.close()
    exp = [l.rstrip("\n") for l in open(os.path.join(G, "collapsed_stitched.records.vcf"))]
    assert len(got) == len(exp)
    for a, b in zip(got, exp):
        assert a == b


@pytest.mark.parametrize("packed", [False, True])
def test_bam_file_to_vcf_text_through_the_library(packed):
    """The whole path from the library's own entry points: PhiX_S3.bam -> pb2_bam_* (decode, read filter) -> pb2_push_reads -> pb2_flush, streamed
    block by block -> pb2_vcf_format == PhiX_S3.noisy.vcf, line by line; and collapsed.test.stitched.bam (XD / XV / XW / XR from the tags)
    -> test_truth.stitched.genome.vcf. packed: the stager hands out pb2_packed_read_batch (one byte per base, compact offsets, tag planes only with
    tagged reads) and the reads go through pb2_push_reads_packed."""
    import ctypes as C
    import os
    pb = _pb()
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    genome = open(os.path.join(G, "phix_genome.txt")).read().strip()
    st = pb.BamReadStager(os.path.join(G, "PhiX_S3.bam"), max_reads=50, packed=packed)
    sm = pb.GpuStateManager(pb.make_config(min_coverage=2, min_base_call_quality=10, min_variant_qscore=1, min_frequency=0.00001, forced_noise_level=40, call_mnvs=1,
                                           max_size_mnv=10, max_gap_mnv=5, no_call_filter=1.0, expect_stitched=int(st.is_stitched)), st.references[0][0], genome)
    caller = pb.GpuAlleleCaller()
    got = []
    for _, batch, _ in st:
        last_pos0 = C.cast(batch.pos0, C.POINTER(C.c_int32))[batch.n_reads - 1]
        sm.AddReadBatch(batch)
        got += sm.FormatVcf(caller.Call(sm, upToPosition=int(last_pos0), raw=True))
    got += sm.FormatVcf(caller.Call(sm, raw=True))
    sm.close()
    st.close()
    assert got == [l.rstrip("\n") for l in open(os.path.join(G, "phix_s3_noisy.records.vcf"))]

    seq = "N" * (9770498 - 1) + ("GAAGTAACAACGCAGGATGCCCCCTGGGGTGGACTGCCCCATGGAATTCTGGACCAAGGAGGAGAATCAGAGCGTTGTGGTTGACTTCCTGCTGCCCACAGGGGTCTACCTGAACTTCCCTGTGTCCCGCAATGCCAACCTC"
                                 "AGCACCATCAAGCAGGTATGGCCTCCATC")
    want = [l.rstrip("\n") for l in open(os.path.join(G, "collapsed_stitched.records.vcf"))]
    for max_reads in (65536, 3, 1):   # streamed in tiny batches too: tagged and untagged reads mix, every batch has the same shape
        st = pb.BamReadStager(os.path.join(G, "collapsed.test.stitched.bam"), max_reads=max_reads, packed=packed)
        sm = pb.GpuStateManager(pb.make_config(call_mnvs=1, max_size_mnv=100, max_gap_mnv=10, expect_stitched=1, expect_collapsed=1, skip_validation=1, amplicon_bias_filter=0.01),
                                "chr1", seq)
        for _, batch, _ in st:
            sm.AddReadBatch(batch)
        recs = pb.GpuAlleleCaller().Call(sm, raw=True)
        got = sm.FormatVcf(recs, sm.AlleleExt(), debug_mode=True, output_bias_files=True, report_rc_counts=True, report_ts_counts=True)
        sm.close()
        st.close()
        assert got == want, max_reads
